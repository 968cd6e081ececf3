// gemm_tc.cu -- node-level dense layers on the 5th-gen tensor cores (tcgen05 + TMEM), fp32-accurate.
//
// Replaces the cuBLAS SGEMM calls of engine.cu (torch.nn.Linear forward / input-gradient of
// nablaDFT/painn_pyg/painn.py:459-464,520-525 and the schnetpack Dense layers).  cuBLAS picks
// a 64x32x16 SIMT tile for these skinny problems (M ~ 10^4 atoms, N,K in {64..384}) and runs at
// ~25 TFLOP/s: 2.1 ms of a 4.6 ms step (profiles/r1_v0_launches.csv).
//
// The reference never uses reduced precision, so single-pass TF32 is
// out.  We use the 3xTF32 split: x = hi + lo with hi = tf32_rn(x), lo = tf32_rn(x - hi);
//   A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi       (dropped term ~2^-24 relative)
// three `tcgen05.mma.kind::tf32` per k-step accumulating in fp32 in TMEM.  The operands are
// split on the fly while they are staged from global into shared memory by the CTA's threads
// (canonical no-swizzle K-major layout: 16-byte k-chunks, rows contiguous), so activations
// never need a pre-pass and weights need no transposed copies (`trans_b` loads B^T directly).
//
//   C[M,N] (ldc) = A[M,K] (lda) . op(B)  (+ C if accumulate)  (+ bias[N])
//   op(B) = B[N,K]^T (ldb, trans_b = 0: Linear forward)  |  B[K,N] (ldb, trans_b = 1: Linear backward)
//   optional second output  act[M,N] = silu(C)   (C then holds the pre-activation)
// Tile: 128 x 64 x 32; one CTA per tile (two resident per SM), 128 threads, double-buffered stages, register prefetch,
// one elected thread issues the MMAs, `tcgen05.commit` -> mbarrier releases a stage.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int G_BM = 128;
constexpr int G_BK = 32;
constexpr int G_THREADS = 512;  // 16 warps: the hi/lo split is SIMT work; with 4 warps per CTA the SM ran at IPC 0.5 (ncu)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): K-major, SWIZZLE_NONE
// start>>4 [0,14) | LBO>>4 [16,30) (stride between 16-byte k-chunks) | SBO>>4 [32,46) (stride
// between 8-row groups) | version=1 [46,48) | layout_type=0 [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32
// [10,13)=2, A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(s_u32(bar)),
        "r"(parity)
        : "memory");
}

// round-to-nearest split (cvt.rna.tf32.f32): |x - hi| <= 2^-12 |x| and the rounding of lo costs
// 2^-24 |x| -- fp32-level and unbiased.  (Masking the low 13 bits instead truncates toward zero:
// measured 2e-6 relative drift of the model energy against the SGEMM path.)
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - hi);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}

// element (row r, k) of a [ROWS x 32] stage tile lives at float index ((k/4)*ROWS + r)*4 + k%4
template <int BN>
struct Stage {
    float a_hi[G_BM * G_BK], a_lo[G_BM * G_BK], b_hi[BN * G_BK], b_lo[BN * G_BK];
};

template <int BN>
__global__ void __launch_bounds__(G_THREADS, 2) k_gemm_tf32x3(int M, int N, int K, const float* __restrict__ A, int lda,
                                                             const float* __restrict__ B, int ldb, int trans_b, float* C, int ldc,
                                                             int accumulate, const float* __restrict__ bias, float* __restrict__ act, int act_kind,
                                                             int batch_kind, long long a_boff, long long b_boff, long long c_boff) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Stage<BN>* stages = reinterpret_cast<Stage<BN>*>(smem_raw);
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem_raw + 2 * sizeof(Stage<BN>));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * G_BM, n0 = blockIdx.y * BN;
    if (batch_kind != 0) {
        // batch over blockIdx.z: A and C advance by a fixed offset; B is selected per batch entry.
        // kind 1 = "lm blocks" of an equivariant feature [rows][(l,m)][channels]: z = (l,m) index,
        // the weight block is W_l (o3.Linear mixes channels per l), bias only on (l,m) = (0,0).
        const int z = blockIdx.z;
        const int bsel = (batch_kind == 1) ? (z >= 16 ? 4 : z >= 9 ? 3 : z >= 4 ? 2 : z >= 1 ? 1 : 0) : z;
        A += (size_t)z * a_boff;
        B += (size_t)bsel * b_boff;
        C += (size_t)z * c_boff;
        if (act) act += (size_t)z * c_boff;
        if (batch_kind == 1 && z > 0) bias = nullptr;
    }

    if (tid == 0) {
        mbar_init_(mma_done, 1);
        mbar_init_(mma_done + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM: 4 accumulators (3 main round-robin + 1 correction) x BN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(4 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    constexpr uint32_t IDESC = umma_idesc_tf32(G_BM, BN);
    constexpr int BPT = BN * (G_BK / 4) / G_THREADS;  // B float4 per thread per chunk (8 for BN=128, 4 for BN=64)
    constexpr int BSPLIT = G_THREADS / BN;            // threads sharing one B row (1 or 2)
    const int n_chunks = K / G_BK;
    constexpr int APT = G_BM * (G_BK / 4) / G_THREADS;  // A float4 per thread per chunk
    const int atile_row = tid % G_BM;                 // this thread stages k-chunks [akc0, akc0 + APT) of row `atile_row` of the A tile
    const int akc0 = (tid / G_BM) * APT;
    const int arow = m0 + atile_row;
    const bool arow_ok = arow < M;
    const int btile_row = tid % BN;                   // and k-chunks [bkc0, bkc0 + BPT) of row `btile_row` of the B tile
    const int bkc0 = (tid / BN) * BPT;
    const int brow = n0 + btile_row;
    const bool brow_ok = brow < N;
    (void)BSPLIT;

    // global -> registers (issued one chunk ahead of the shared-memory stores: the L2 latency of
    // chunk c+1 overlaps the split/store/MMA of chunk c)
    auto gload = [&](int ch, float4 (&ra)[APT], float4 (&rb)[BPT]) {
        const int k0 = ch * G_BK;
        const float* src = A + (size_t)arow * lda + k0;
#pragma unroll
        for (int i = 0; i < APT; ++i) ra[i] = arow_ok ? ldg4(src + 4 * (akc0 + i)) : f4(0.f);
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int kc = bkc0 + i;
            float4 v = f4(0.f);
            if (brow_ok) {
                if (!trans_b) {
                    v = ldg4(B + (size_t)brow * ldb + k0 + 4 * kc);
                } else {  // B[k][n]: four k-rows, coalesced across the threads of a warp (consecutive n)
                    const float* p = B + (size_t)(k0 + 4 * kc) * ldb + brow;
                    v = make_float4(__ldg(p), __ldg(p + ldb), __ldg(p + 2 * (size_t)ldb), __ldg(p + 3 * (size_t)ldb));
                }
            }
            rb[i] = v;
        }
    };
    // registers -> hi/lo split -> shared ([k-chunk][row] 16-byte units: conflict-free), then MMAs
    auto process = [&](int ch, const float4 (&ra)[APT], const float4 (&rb)[BPT]) {
        const int s = ch & 1, use = ch >> 1;
        if (use > 0) mbar_wait_(mma_done + s, (uint32_t)((use - 1) & 1));  // MMAs that read this buffer are done
        Stage<BN>& st = stages[s];
#pragma unroll
        for (int i = 0; i < APT; ++i) {
            float4 hi, lo;
            split4(ra[i], hi, lo);
            st4(st.a_hi + ((akc0 + i) * G_BM + atile_row) * 4, hi);
            st4(st.a_lo + ((akc0 + i) * G_BM + atile_row) * 4, lo);
        }
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            float4 hi, lo;
            split4(rb[i], hi, lo);
            st4(st.b_hi + ((bkc0 + i) * BN + btile_row) * 4, hi);
            st4(st.b_lo + ((bkc0 + i) * BN + btile_row) * 4, lo);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint64_t dah = umma_desc(s_u32(st.a_hi), G_BM * 16, 128), dal = umma_desc(s_u32(st.a_lo), G_BM * 16, 128);
            uint64_t dbh = umma_desc(s_u32(st.b_hi), BN * 16, 128), dbl = umma_desc(s_u32(st.b_lo), BN * 16, 128);
#pragma unroll
            for (int ks = 0; ks < G_BK / 8; ++ks) {  // one MMA = 8 k-values = two 16-byte chunks
                // The tensor core truncates when it adds into the fp32 accumulator: the error grows
                // linearly with the chain length (measured 7.6e-9 * K relative with one accumulator).
                // So: the O(2^-11) correction terms get their own accumulator, and the main term
                // rotates over three accumulators; the four are summed with RN adds in the epilogue.
                const int g = ch * (G_BK / 8) + ks;
                umma_tf32(tmem_acc + 3 * BN, dal, dbh, IDESC, g > 0 ? 1u : 0u);
                umma_tf32(tmem_acc + 3 * BN, dah, dbl, IDESC, 1u);
                umma_tf32(tmem_acc + (g % 3) * BN, dah, dbh, IDESC, g >= 3 ? 1u : 0u);
                dah += (2 * G_BM * 16) >> 4; dal += (2 * G_BM * 16) >> 4;  // start-address field counts 16-byte units
                dbh += (2 * BN * 16) >> 4; dbl += (2 * BN * 16) >> 4;
            }
            // arrives on the mbarrier when every MMA issued so far has completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(mma_done + s)) : "memory");
        }
    };

    {
        float4 ra0[APT], rb0[BPT], ra1[APT], rb1[BPT];
        gload(0, ra0, rb0);
        for (int ch = 0; ch < n_chunks; ch += 2) {
            if (ch + 1 < n_chunks) gload(ch + 1, ra1, rb1);
            process(ch, ra0, rb0);
            if (ch + 1 < n_chunks) {
                if (ch + 2 < n_chunks) gload(ch + 2, ra0, rb0);
                process(ch + 1, ra1, rb1);
            }
        }
    }
    {   // accumulator complete when the last chunk's commit has arrived (MMAs retire in order)
        const int last = n_chunks - 1;
        mbar_wait_(mma_done + (last & 1), (uint32_t)((last >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- epilogue: warp w owns TMEM lanes [32w, 32w+32) == tile rows; thread = one row, 32 columns per load
    // warp w may touch TMEM lanes [32 (w % 4), +32) only; the column range is split over the warp quads
    constexpr int COLS_PER_WARP = BN / (G_THREADS / 128);
    const int lane_grp = warp & 3;
    const int row = m0 + lane_grp * 32 + (tid & 31);
#pragma unroll 1
    for (int cb = (warp >> 2) * COLS_PER_WARP; cb < ((warp >> 2) + 1) * COLS_PER_WARP; cb += 16) {
        uint32_t r[4][16];
#pragma unroll
        for (int acc = 0; acc < 4; ++acc) {
            const uint32_t taddr = tmem_acc + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * BN + cb);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[acc][0]), "=r"(r[acc][1]), "=r"(r[acc][2]), "=r"(r[acc][3]), "=r"(r[acc][4]), "=r"(r[acc][5]), "=r"(r[acc][6]),
                  "=r"(r[acc][7]), "=r"(r[acc][8]), "=r"(r[acc][9]), "=r"(r[acc][10]), "=r"(r[acc][11]), "=r"(r[acc][12]), "=r"(r[acc][13]),
                  "=r"(r[acc][14]), "=r"(r[acc][15])
                : "r"(taddr)
                : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float v16[16];
#pragma unroll
        for (int t = 0; t < 16; ++t)
            v16[t] = (__uint_as_float(r[0][t]) + __uint_as_float(r[1][t])) + (__uint_as_float(r[2][t]) + __uint_as_float(r[3][t]));
        if (row < M) {
            const int nb = n0 + cb;
            float* crow = C + (size_t)row * ldc + nb;
            float* arow_out = act ? act + (size_t)row * ldc + nb : nullptr;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                if (nb + 4 * q4 < N) {
                    float4 v = make_float4(v16[4 * q4], v16[4 * q4 + 1], v16[4 * q4 + 2], v16[4 * q4 + 3]);
                    if (bias) v = v + ldg4(bias + nb + 4 * q4);
                    if (accumulate) v = v + *reinterpret_cast<const float4*>(crow + 4 * q4);
                    st4(crow + 4 * q4, v);
                    if (arow_out) st4(arow_out + 4 * q4, make_float4(actf_(v.x, act_kind), actf_(v.y, act_kind), actf_(v.z, act_kind), actf_(v.w, act_kind)));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(4 * BN) : "memory");
}

template <int BN>
int launch(int M, int N, int K, const float* A, int lda, const float* B, int ldb, int trans_b, float* C, int ldc, int accumulate,
           const float* bias, float* act, int act_kind, int batch_kind, int n_batch, long long a_boff, long long b_boff, long long c_boff,
           cudaStream_t s) {
    const int smem = 2 * (int)sizeof(Stage<BN>) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_gemm_tf32x3<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return nb_check_launch();
        attr_set = true;
    }
    dim3 grid((M + G_BM - 1) / G_BM, (N + BN - 1) / BN, batch_kind ? n_batch : 1);
    k_gemm_tf32x3<BN><<<grid, G_THREADS, smem, s>>>(M, N, K, A, lda, B, ldb, trans_b, C, ldc, accumulate, bias, act, act_kind, batch_kind, a_boff,
                                                    b_boff, c_boff);
    return nb_check_launch();
}

// ---------------------------------------------------------------------------------------------------------
// A-stationary variant for wide outputs (N >= 512, K <= 128): QHNet's weight-generation layers are
// [P ~ 10^5, 128] x [128, 8320].  With the tile kernel above every 64-column tile re-stages and re-splits its
// 128 x K slab of A (130 times for N = 8320) and the GEMM runs at 46 TFLOP/s (tools/gemm_microbench.py).
// Here a CTA stages + splits its A slab ONCE (hi/lo, 2 x 64 KB), then walks the N dimension in 32-column tiles:
// B tiles are double-buffered in shared memory, accumulators are double-buffered in TMEM (2 x 4 x 32 columns), so
// the tensor core works on tile t while the CTA's threads drain tile t-1 (TMEM -> registers -> global) and stage
// tile t+1.  Same 3xTF32 split and 3+1 accumulator rotation as above.
constexpr int AS_BN = 32;
constexpr int AS_KMAX = 128;

__global__ void __launch_bounds__(G_THREADS, 1) k_gemm_tf32x3_as(int M, int N, int K, const float* __restrict__ A, int lda,
                                                                const float* __restrict__ B, int ldb, int trans_b, float* __restrict__ C, int ldc,
                                                                const float* __restrict__ bias, float* __restrict__ act, int act_kind,
                                                                int tiles_per_cta) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* a_hi = reinterpret_cast<float*>(smem_raw);
    float* a_lo = a_hi + G_BM * AS_KMAX;
    float* b_buf = a_lo + G_BM * AS_KMAX;  // [2][hi|lo][AS_BN * AS_KMAX]
    uint64_t* acc_done = reinterpret_cast<uint64_t*>(b_buf + 4 * AS_BN * AS_KMAX);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * G_BM;
    const int n_tiles_total = (N + AS_BN - 1) / AS_BN;
    const int t_begin = blockIdx.y * tiles_per_cta, t_end = min(t_begin + tiles_per_cta, n_tiles_total);
    if (t_begin >= t_end) return;

    if (tid == 0) {
        mbar_init_(acc_done, 1);
        mbar_init_(acc_done + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(8 * AS_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- stage + split the whole A slab once: thread -> row tid % 128, k-chunks (tid / 128) + 4 j
    {
        const int row = tid % G_BM, kc0 = tid / G_BM;
        const bool ok = m0 + row < M;
        const float* src = A + (size_t)(m0 + row) * lda;
        for (int kc = kc0; kc < K / 4; kc += G_THREADS / G_BM) {
            const float4 v = ok ? ldg4(src + 4 * kc) : f4(0.f);
            float4 hi, lo;
            split4(v, hi, lo);
            st4(a_hi + (kc * G_BM + row) * 4, hi);
            st4(a_lo + (kc * G_BM + row) * 4, lo);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t IDESC = umma_idesc_tf32(G_BM, AS_BN);
    const int kchunks = K / 4;                 // 16-byte k-chunks
    const int b_items = AS_BN * kchunks;       // float4 per B tile
    const int lane_grp = warp & 3, colgrp = warp >> 2;
    const int row = m0 + lane_grp * 32 + (tid & 31);

    auto load_b = [&](int t, float4 (&rb)[2]) {  // global -> registers (<= 2 float4 per thread: 32 x 128 / 4 / 512)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int item = tid + i * G_THREADS;
            float4 v = f4(0.f);
            if (item < b_items) {
                const int n = item % AS_BN, kc = item / AS_BN;
                const int gn = t * AS_BN + n;
                if (gn < N) {
                    if (!trans_b) v = ldg4(B + (size_t)gn * ldb + 4 * kc);
                    else {
                        const float* p = B + (size_t)(4 * kc) * ldb + gn;
                        v = make_float4(__ldg(p), __ldg(p + ldb), __ldg(p + 2 * (size_t)ldb), __ldg(p + 3 * (size_t)ldb));
                    }
                }
            }
            rb[i] = v;
        }
    };
    auto store_b = [&](int buf, const float4 (&rb)[2]) {
        float* bh = b_buf + buf * 2 * AS_BN * AS_KMAX;
        float* bl = bh + AS_BN * AS_KMAX;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int item = tid + i * G_THREADS;
            if (item < b_items) {
                const int n = item % AS_BN, kc = item / AS_BN;
                float4 hi, lo;
                split4(rb[i], hi, lo);
                st4(bh + (kc * AS_BN + n) * 4, hi);
                st4(bl + (kc * AS_BN + n) * 4, lo);
            }
        }
    };
    auto drain = [&](int t, int buf) {  // epilogue of tile t from accumulator set `buf`
        const int it = t - t_begin;
        mbar_wait_(acc_done + buf, (uint32_t)((it >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // all four accumulators are fetched back to back and waited on once (four dependent ld+wait round
        // trips were ~40 % of the per-tile time)
        uint32_t r[4][8];
#pragma unroll
        for (int acc = 0; acc < 4; ++acc) {
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(buf * 4 * AS_BN + acc * AS_BN + colgrp * 8);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[acc][0]), "=r"(r[acc][1]), "=r"(r[acc][2]), "=r"(r[acc][3]), "=r"(r[acc][4]), "=r"(r[acc][5]), "=r"(r[acc][6]),
                           "=r"(r[acc][7])
                         : "r"(taddr)
                         : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float v8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
            v8[q] = (__uint_as_float(r[0][q]) + __uint_as_float(r[1][q])) + (__uint_as_float(r[2][q]) + __uint_as_float(r[3][q]));
        if (row < M) {
            const int nb = t * AS_BN + colgrp * 8;
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
                if (nb + 4 * q4 < N) {
                    float4 v = make_float4(v8[4 * q4], v8[4 * q4 + 1], v8[4 * q4 + 2], v8[4 * q4 + 3]);
                    if (bias) v = v + ldg4(bias + nb + 4 * q4);
                    if (act) st4(act + (size_t)row * ldc + nb + 4 * q4, make_float4(actf_(v.x, act_kind), actf_(v.y, act_kind), actf_(v.z, act_kind), actf_(v.w, act_kind)));
                    else st4(C + (size_t)row * ldc + nb + 4 * q4, v);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    float4 rb[2];
    load_b(t_begin, rb);
    for (int t = t_begin; t < t_end; ++t) {
        const int it = t - t_begin, buf = it & 1;
        // B buffer `buf` was last read by the MMAs of tile t-2, whose completion was awaited in drain(t-2)
        store_b(buf, rb);
        if (t + 1 < t_end) load_b(t + 1, rb);  // prefetch the next tile's rows while this one is multiplied
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();  // also orders drain(t-2)'s TMEM reads of accumulator set `buf` before the MMAs below
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // descriptors differ between k-steps only in the start-address field (bits 0-13, units of 16 B):
            // build them once and step with one 64-bit add -- the single issuing thread must sustain one MMA
            // per ~16 cycles (128 x 32 x 8 tile), which the per-MMA descriptor rebuild could not.
            const uint32_t bh = s_u32(b_buf + buf * 2 * AS_BN * AS_KMAX), bl = bh + AS_BN * AS_KMAX * 4;
            uint64_t dah = umma_desc(s_u32(a_hi), G_BM * 16, 128), dal = umma_desc(s_u32(a_lo), G_BM * 16, 128);
            uint64_t dbh = umma_desc(bh, AS_BN * 16, 128), dbl = umma_desc(bl, AS_BN * 16, 128);
            const uint32_t acc0 = tmem_base + buf * 4 * AS_BN;
            const int nks = K / 8;
#pragma unroll 4
            for (int ks = 0; ks < nks; ++ks) {
                umma_tf32(acc0 + 3 * AS_BN, dal, dbh, IDESC, ks > 0 ? 1u : 0u);
                umma_tf32(acc0 + 3 * AS_BN, dah, dbl, IDESC, 1u);
                umma_tf32(acc0 + (ks % 3) * AS_BN, dah, dbh, IDESC, ks >= 3 ? 1u : 0u);
                dah += (2 * G_BM * 16) >> 4; dal += (2 * G_BM * 16) >> 4;
                dbh += (2 * AS_BN * 16) >> 4; dbl += (2 * AS_BN * 16) >> 4;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(acc_done + buf)) : "memory");
        }
        if (t > t_begin) drain(t - 1, buf ^ 1);  // overlaps with the tensor core working on tile t
    }
    drain(t_end - 1, (t_end - 1 - t_begin) & 1);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(8 * AS_BN) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// TS variant: the A slab lives in TENSOR MEMORY (tcgen05.mma with the A operand from TMEM).
// Measured on the SS kernels above: a 128 x 32 x 8 MMA reads 4 KB of A + 1 KB of B from shared memory per 16
// cycles = 320 B/clk against the 128 B/clk shared-memory port, so the tensor pipe starves (59 TFLOP/s on
// [10^5,128]x[128,8320]).  Here each thread splits its row of A once and parks hi / lo in TMEM columns
// [0,128) / [128,256) with tcgen05.st; the MMAs then read only the 32 x 8 B tile from shared memory (64 B/clk).
// TMEM budget (512 columns): A_hi 128 | A_lo 128 | 2 accumulator sets x (3 main + 1 correction) x 32.
// C = A.op(B) (+C) (+bias); act (optional) = activation(C).  K <= 128 per launch (longer K: chained launches).
constexpr int TS_BN = 32;
constexpr int TS_A_HI = 0, TS_A_LO = 128, TS_ACC = 256;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__global__ void __launch_bounds__(G_THREADS, 1) k_gemm_tf32x3_ts(int M, int N, int K, const float* __restrict__ A, int lda,
                                                                const float* __restrict__ B, int ldb, int trans_b, float* C, int ldc, int accumulate,
                                                                const float* __restrict__ bias, float* __restrict__ act, int act_kind,
                                                                int tiles_per_cta) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* b_buf = reinterpret_cast<float*>(smem_raw);  // [2][hi|lo][TS_BN * 128]
    uint64_t* acc_done = reinterpret_cast<uint64_t*>(b_buf + 4 * TS_BN * AS_KMAX);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * G_BM;
    const int n_tiles_total = (N + TS_BN - 1) / TS_BN;
    const int t_begin = blockIdx.y * tiles_per_cta, t_end = min(t_begin + tiles_per_cta, n_tiles_total);
    if (t_begin >= t_end) return;

    if (tid == 0) {
        mbar_init_(acc_done, 1);
        mbar_init_(acc_done + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int lane_grp = warp & 3, colgrp = warp >> 2;
    const int row = m0 + lane_grp * 32 + lane;
    // ---- A slab -> TMEM: warp (lane_grp, colgrp) owns rows [32 lane_grp, +32) x k in [32 colgrp, +32)
    if (32 * colgrp < K) {
        uint32_t hi[32], lo[32];
        const float* src = A + (size_t)row * lda + 32 * colgrp;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = (row < M) ? ldg4(src + 4 * q) : f4(0.f);
            float4 h, l;
            split4(v, h, l);
            hi[4 * q] = __float_as_uint(h.x); hi[4 * q + 1] = __float_as_uint(h.y); hi[4 * q + 2] = __float_as_uint(h.z); hi[4 * q + 3] = __float_as_uint(h.w);
            lo[4 * q] = __float_as_uint(l.x); lo[4 * q + 1] = __float_as_uint(l.y); lo[4 * q + 2] = __float_as_uint(l.z); lo[4 * q + 3] = __float_as_uint(l.w);
        }
        const uint32_t ta = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(32 * colgrp);
#define NB_TMEM_ST32(ADDR, R)                                                                                                              \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22," \
                 "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(ADDR),                                                                  \
                 "r"(R[0]), "r"(R[1]), "r"(R[2]), "r"(R[3]), "r"(R[4]), "r"(R[5]), "r"(R[6]), "r"(R[7]), "r"(R[8]), "r"(R[9]), "r"(R[10]),      \
                 "r"(R[11]), "r"(R[12]), "r"(R[13]), "r"(R[14]), "r"(R[15]), "r"(R[16]), "r"(R[17]), "r"(R[18]), "r"(R[19]), "r"(R[20]),       \
                 "r"(R[21]), "r"(R[22]), "r"(R[23]), "r"(R[24]), "r"(R[25]), "r"(R[26]), "r"(R[27]), "r"(R[28]), "r"(R[29]), "r"(R[30]),       \
                 "r"(R[31])                                                                                                                   \
                 : "memory")
        NB_TMEM_ST32(ta + TS_A_HI, hi);
        NB_TMEM_ST32(ta + TS_A_LO, lo);
#undef NB_TMEM_ST32
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    constexpr uint32_t IDESC = umma_idesc_tf32(G_BM, TS_BN);
    const int kchunks = K / 4;
    const int b_items = TS_BN * kchunks;

    auto load_b = [&](int t, float4 (&rb)[2]) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int item = tid + i * G_THREADS;
            float4 v = f4(0.f);
            if (item < b_items) {
                const int n = item % TS_BN, kc = item / TS_BN;
                const int gn = t * TS_BN + n;
                if (gn < N) {
                    if (!trans_b) v = ldg4(B + (size_t)gn * ldb + 4 * kc);
                    else {
                        const float* p = B + (size_t)(4 * kc) * ldb + gn;
                        v = make_float4(__ldg(p), __ldg(p + ldb), __ldg(p + 2 * (size_t)ldb), __ldg(p + 3 * (size_t)ldb));
                    }
                }
            }
            rb[i] = v;
        }
    };
    auto store_b = [&](int buf, const float4 (&rb)[2]) {
        float* bh = b_buf + buf * 2 * TS_BN * AS_KMAX;
        float* bl = bh + TS_BN * AS_KMAX;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int item = tid + i * G_THREADS;
            if (item < b_items) {
                const int n = item % TS_BN, kc = item / TS_BN;
                float4 hi, lo;
                split4(rb[i], hi, lo);
                st4(bh + (kc * TS_BN + n) * 4, hi);
                st4(bl + (kc * TS_BN + n) * 4, lo);
            }
        }
    };
    auto drain = [&](int t, int buf) {
        const int it = t - t_begin;
        mbar_wait_(acc_done + buf, (uint32_t)((it >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[4][8];
#pragma unroll
        for (int acc = 0; acc < 4; ++acc) {
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(TS_ACC + buf * 4 * TS_BN + acc * TS_BN + colgrp * 8);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[acc][0]), "=r"(r[acc][1]), "=r"(r[acc][2]), "=r"(r[acc][3]), "=r"(r[acc][4]), "=r"(r[acc][5]), "=r"(r[acc][6]),
                           "=r"(r[acc][7])
                         : "r"(taddr)
                         : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M) {
            const int nb = t * TS_BN + colgrp * 8;
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
                if (nb + 4 * q4 < N) {
                    float4 v;
                    v.x = (__uint_as_float(r[0][4 * q4]) + __uint_as_float(r[1][4 * q4])) + (__uint_as_float(r[2][4 * q4]) + __uint_as_float(r[3][4 * q4]));
                    v.y = (__uint_as_float(r[0][4 * q4 + 1]) + __uint_as_float(r[1][4 * q4 + 1])) + (__uint_as_float(r[2][4 * q4 + 1]) + __uint_as_float(r[3][4 * q4 + 1]));
                    v.z = (__uint_as_float(r[0][4 * q4 + 2]) + __uint_as_float(r[1][4 * q4 + 2])) + (__uint_as_float(r[2][4 * q4 + 2]) + __uint_as_float(r[3][4 * q4 + 2]));
                    v.w = (__uint_as_float(r[0][4 * q4 + 3]) + __uint_as_float(r[1][4 * q4 + 3])) + (__uint_as_float(r[2][4 * q4 + 3]) + __uint_as_float(r[3][4 * q4 + 3]));
                    float* cp = C + (size_t)row * ldc + nb + 4 * q4;
                    if (bias) v = v + ldg4(bias + nb + 4 * q4);
                    if (accumulate) v = v + *reinterpret_cast<const float4*>(cp);
                    st4(cp, v);
                    if (act) st4(act + (size_t)row * ldc + nb + 4 * q4, make_float4(actf_(v.x, act_kind), actf_(v.y, act_kind), actf_(v.z, act_kind), actf_(v.w, act_kind)));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    float4 rb[2];
    load_b(t_begin, rb);
    for (int t = t_begin; t < t_end; ++t) {
        const int it = t - t_begin, buf = it & 1;
        store_b(buf, rb);
        if (t + 1 < t_end) load_b(t + 1, rb);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bh = s_u32(b_buf + buf * 2 * TS_BN * AS_KMAX), bl = bh + TS_BN * AS_KMAX * 4;
            uint64_t dbh = umma_desc(bh, TS_BN * 16, 128), dbl = umma_desc(bl, TS_BN * 16, 128);
            const uint32_t acc0 = tmem_base + TS_ACC + buf * 4 * TS_BN;
            uint32_t ah = tmem_base + TS_A_HI, al = tmem_base + TS_A_LO;
            const int nks = K / 8;
#pragma unroll 4
            for (int ks = 0; ks < nks; ++ks) {
                umma_tf32_ts(acc0 + 3 * TS_BN, al, dbh, IDESC, ks > 0 ? 1u : 0u);
                umma_tf32_ts(acc0 + 3 * TS_BN, ah, dbl, IDESC, 1u);
                umma_tf32_ts(acc0 + (ks % 3) * TS_BN, ah, dbh, IDESC, ks >= 3 ? 1u : 0u);
                ah += 8; al += 8;  // 8 tf32 k-values = 8 TMEM columns
                dbh += (2 * TS_BN * 16) >> 4; dbl += (2 * TS_BN * 16) >> 4;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(acc_done + buf)) : "memory");
        }
        if (t > t_begin) drain(t - 1, buf ^ 1);
    }
    drain(t_end - 1, (t_end - 1 - t_begin) & 1);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

int launch_ts(int M, int N, int K, const float* A, int lda, const float* B, int ldb, int trans_b, float* C, int ldc, int accumulate,
              const float* bias, float* act, int act_kind, cudaStream_t s) {
    const int smem = 4 * TS_BN * AS_KMAX * (int)sizeof(float) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_gemm_tf32x3_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return nb_check_launch();
        attr_set = true;
    }
    const int m_tiles = (M + G_BM - 1) / G_BM, n_tiles = (N + TS_BN - 1) / TS_BN;
    int ny = 1;  // split the N walk only when there are too few row slabs to fill the 148 SMs
    while (m_tiles * ny < 148 && ny * 2 <= n_tiles) ny *= 2;
    const int tiles_per_cta = (n_tiles + ny - 1) / ny;
    dim3 grid(m_tiles, (n_tiles + tiles_per_cta - 1) / tiles_per_cta);
    k_gemm_tf32x3_ts<<<grid, G_THREADS, smem, s>>>(M, N, K, A, lda, B, ldb, trans_b, C, ldc, accumulate, bias, act, act_kind, tiles_per_cta);
    return nb_check_launch();
}

int launch_as(int M, int N, int K, const float* A, int lda, const float* B, int ldb, int trans_b, float* C, int ldc, const float* bias,
              float* act, int act_kind, cudaStream_t s) {
    const int smem = (2 * G_BM * AS_KMAX + 4 * AS_BN * AS_KMAX) * (int)sizeof(float) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_gemm_tf32x3_as, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return nb_check_launch();
        attr_set = true;
    }
    const int m_tiles = (M + G_BM - 1) / G_BM, n_tiles = (N + AS_BN - 1) / AS_BN;
    int ny = 1;  // split the N walk only when there are too few row slabs to fill the 148 SMs
    while (m_tiles * ny < 148 && ny < n_tiles / 4) ny *= 2;
    const int tiles_per_cta = (n_tiles + ny - 1) / ny;
    dim3 grid(m_tiles, (n_tiles + tiles_per_cta - 1) / tiles_per_cta);
    k_gemm_tf32x3_as<<<grid, G_THREADS, smem, s>>>(M, N, K, A, lda, B, ldb, trans_b, C, ldc, bias, act, act_kind, tiles_per_cta);
    return nb_check_launch();
}

}  // namespace

// Constraints: K % 32 == 0, N % 4 == 0, lda/ldb/ldc % 4 == 0, 16-byte aligned pointers; `act` (optional) shares ldc with C.
int nb_gemm_tf32x3_ex(int M, int N, int K, const float* A, int lda, const float* B, int ldb, int trans_b, float* C, int ldc, int accumulate,
                      const float* bias, float* act, int act_kind, cudaStream_t s) {
    if (!A || !B || !C || M < 0 || N <= 0 || K <= 0) return NB200_EINVAL;
    if (K % G_BK || N % 4 || lda % 4 || ldb % 4 || ldc % 4) return NB200_EUNSUPPORTED;
    if (M == 0) return NB200_OK;
    static const int variant = [] { const char* e = getenv("NB200_GEMM_VARIANT"); return !e ? 2 : (e[0] == 't' && e[1] == 'i') ? 0 : (e[0] == 'a') ? 1 : 2; }();
    if (variant == 2 && N >= 32) {
        // A-in-TMEM kernel; K > 128 is chained in 128-wide launches that accumulate into C (bias first, activation last)
        for (int k0 = 0; k0 < K; k0 += AS_KMAX) {
            const int kc = (K - k0 < AS_KMAX) ? (K - k0) : AS_KMAX;
            const bool last = k0 + kc >= K;
            const float* Bk = trans_b ? B + (size_t)k0 * ldb : B + k0;
            int rc = launch_ts(M, N, kc, A + k0, lda, Bk, ldb, trans_b, C, ldc, (accumulate || k0 > 0) ? 1 : 0, k0 == 0 ? bias : nullptr,
                               last ? act : nullptr, act_kind, s);
            if (rc != NB200_OK) return rc;
        }
        return NB200_OK;
    }
    if (variant >= 1 && N >= 512 && K <= AS_KMAX && !accumulate && M >= 1024) {
        // `act` requested: the A-stationary epilogue writes only the activation (callers that need the
        // pre-activation too -- the PaiNN backward -- have N <= 384 and never come here)
        return launch_as(M, N, K, A, lda, B, ldb, trans_b, C, ldc, bias, act, act_kind, s);
    }
    // BN = 64: 4 x 64 TMEM columns and 96 KB of stages per CTA -> two CTAs per SM and twice as many
    // tiles, which matters more than tile efficiency for these skinny (M ~ 10^4, N <= 384) problems
    return launch<64>(M, N, K, A, lda, B, ldb, trans_b, C, ldc, accumulate, bias, act, act_kind, 0, 1, 0, 0, 0, s);
}

// batched over the 25 (l,m) rows of an equivariant feature: see batch_kind 1 in the kernel
int nb_gemm_tf32x3_lm(int M, int N, int K, const float* A, int lda, const float* W_l, long long w_l_stride, float* C, int ldc, int accumulate,
                      const float* bias, int n_lm, cudaStream_t s) {
    if (!A || !W_l || !C || M < 0 || N <= 0 || K <= 0) return NB200_EINVAL;
    if (K % G_BK || N % 4 || lda % 4 || ldc % 4) return NB200_EUNSUPPORTED;
    if (M == 0) return NB200_OK;
    return launch<64>(M, N, K, A, lda, W_l, N, 1, C, ldc, accumulate, bias, nullptr, NB_ACT_SILU, 1, n_lm, K, w_l_stride, N, s);
}

extern "C" int nb200_gemm_tf32x3(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* B, int32_t ldb,
                                 int32_t trans_b, float* C, int32_t ldc, int32_t accumulate, const float* bias, float* act,
                                 void* stream) {
    return nb_gemm_tf32x3_ex(M, N, K, A, lda, B, ldb, trans_b, C, ldc, accumulate, bias, act, NB_ACT_SILU, (cudaStream_t)stream);
}
