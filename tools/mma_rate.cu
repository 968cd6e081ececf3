// mma_rate.cu -- microbenchmark behind the fused node kernel's design.
//
// Question: for the fp32-accurate 3xTF32 scheme (three tcgen05.mma.kind::tf32 M128 N128 K8 per k-step), what bounds the issue rate
// on B200 -- the tensor pipe (64 cycles per MMA) or operand fetch from shared memory -- and does parking the A operand in
// tensor memory (TS form) move it?  Also: how fast do 8 warps drain accumulators with tcgen05.ld?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu && tools/mma_rate
//
// Modes (cycles per MMA reported, averaged over CTAs; `bg` = a second warp streams cp.async.bulk copies into shared memory meanwhile):
//   0  SS     A_hi, A_lo, B_hi, B_lo all from shared memory
//   1  hybrid A_hi from TMEM, A_lo from shared memory
//   2  TS     A_hi and A_lo from TMEM
//   6  tcgen05.ld drain: 8 warps read 3 x [128 x 128] fp32 accumulators
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
            s_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)), "l"(src), "r"(bytes),
                 "r"(s_u32(bar))
                 : "memory");
}

constexpr int LBO = 128 * 16 + 16;          // bytes between 16-byte k-chunks (padded, as gemm_tc.cu)
constexpr int A_BYTES = 32 * LBO;           // one of hi / lo, K = 128
constexpr int BST_BYTES = 2 * 8 * LBO;      // one ring stage: hi + lo, 32 k
constexpr int SMEM = 2 * A_BYTES + 3 * BST_BYTES + 256;

__global__ void __launch_bounds__(320, 1) k_rate(int mode, int bg, int tiles, const float* gsrc, unsigned long long* out_cyc,
                                                 unsigned long long* out_bg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* a_hi = reinterpret_cast<float*>(smem);
    float* a_lo = reinterpret_cast<float*>(smem + A_BYTES);
    unsigned char* ring = smem + 2 * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * A_BYTES + 3 * BST_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    volatile int* stop = reinterpret_cast<volatile int*>(tmem_slot + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < (2 * A_BYTES + 3 * BST_BYTES) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f + (float)(i & 7) * 0.125f;
    if (tid == 0) {
        for (int b = 0; b < 8; ++b) mbar_init(bars + b, 1);
        *stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *tmem_slot;
    if (warp < 4) {  // fill TMEM columns [0,256) with finite values (A operand when read in TS form)
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(1.0f + 0.125f * (float)(i & 7));
        for (int c = 0; c < 256; c += 32) {
            const uint32_t ta = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
                "%26,%27,%28,%29,%30,%31,%32};" ::"r"(ta),
                "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
                "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    constexpr uint32_t IDESC = idesc_tf32(128, 128);
    if (mode <= 2) {
        if (warp == 8 && lane == 0) {
            const uint64_t da_hi0 = umma_desc(s_u32(a_hi), LBO, 128), da_lo0 = umma_desc(s_u32(a_lo), LBO, 128);
            const long long t0 = clock64();
            for (int t = 0; t < tiles; ++t) {
                uint64_t da_hi = da_hi0, da_lo = da_lo0;
                uint32_t ta_hi = tm, ta_lo = tm + 128;
                for (int ks = 0; ks < 16; ++ks) {
                    const int st = (ks >> 2) & 1;  // ring stages 0 and 1 only (stage 2 is the background-copy target)
                    const uint32_t bh = s_u32(ring + st * BST_BYTES) + (uint32_t)((ks & 3) * 2 * LBO);
                    const uint64_t db_hi = umma_desc(bh, LBO, 128), db_lo = umma_desc(bh + 8 * LBO, LBO, 128);
                    if (mode == 0) {
                        mma_ss(tm + 128, da_lo, db_hi, IDESC, ks > 0);
                        mma_ss(tm + 128, da_hi, db_lo, IDESC, 1);
                        mma_ss(tm + 256 + (ks & 1) * 128, da_hi, db_hi, IDESC, ks >= 2);
                    } else if (mode == 1) {
                        mma_ss(tm + 128, da_lo, db_hi, IDESC, ks > 0);
                        mma_ts(tm + 128, ta_hi, db_lo, IDESC, 1);
                        mma_ts(tm + 256 + (ks & 1) * 128, ta_hi, db_hi, IDESC, ks >= 2);
                    } else {
                        mma_ts(tm + 256, ta_lo, db_hi, IDESC, ks > 0);
                        mma_ts(tm + 256, ta_hi, db_lo, IDESC, 1);
                        mma_ts(tm + 384, ta_hi, db_hi, IDESC, ks >= 1);
                    }
                    da_hi += (2 * LBO) >> 4; da_lo += (2 * LBO) >> 4;
                    ta_hi += 8; ta_lo += 8;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bars)) : "memory");
            mbar_wait(bars, 0);
            const long long t1 = clock64();
            *stop = 1;
            out_cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
        } else if (warp == 9 && lane == 0 && bg) {
            // background: two bulk copies of one ring stage in flight, back to back, into ring stage 2
            unsigned long long n = 0;
            uint32_t ph[2] = {0, 0};
            unsigned char* dst = ring + 2 * BST_BYTES;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(gsrc) + (size_t)(blockIdx.x % 64) * BST_BYTES;
            for (int b = 0; b < 2; ++b) { mbar_expect_tx(bars + 2 + b, BST_BYTES / 2); bulk_g2s(dst + b * (BST_BYTES / 2), src + b * (BST_BYTES / 2), BST_BYTES / 2, bars + 2 + b); }
            while (!*stop) {
                for (int b = 0; b < 2; ++b) {
                    mbar_wait(bars + 2 + b, ph[b]); ph[b] ^= 1; ++n;
                    mbar_expect_tx(bars + 2 + b, BST_BYTES / 2);
                    bulk_g2s(dst + b * (BST_BYTES / 2), src + b * (BST_BYTES / 2), BST_BYTES / 2, bars + 2 + b);
                }
            }
            for (int b = 0; b < 2; ++b) mbar_wait(bars + 2 + b, ph[b]);
            out_bg[blockIdx.x] = n * (BST_BYTES / 2);
        }
    } else if (mode == 6) {
        __syncthreads();
        if (warp < 8) {
            const int lane_grp = warp & 3, chalf = warp >> 2;
            float acc = 0.f;
            const long long t0 = clock64();
            for (int t = 0; t < tiles; ++t) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t r[32];
                        const uint32_t ta = tm + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(128 + j * 128 + chalf * 64 + h * 32);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
                            "%25,%26,%27,%28,%29,%30,%31}, [%32];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                            : "r"(ta)
                            : "memory");
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
                    }
                }
            }
            const long long t1 = clock64();
            if (acc == 123.456f) out_bg[0] = 1;
            if (tid == 0) out_cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

int main(int argc, char** argv) {
    const int tiles = argc > 1 ? atoi(argv[1]) : 40;
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    float* gsrc;
    CK(cudaMalloc(&gsrc, 64 * BST_BYTES));
    CK(cudaMemset(gsrc, 0, 64 * BST_BYTES));
    unsigned long long *d_cyc, *d_bg;
    CK(cudaMalloc(&d_cyc, 148 * 8));
    CK(cudaMalloc(&d_bg, 148 * 8));
    printf("smem per CTA %d B, %d tiles of 16 k-steps x 3 MMAs (M128 N128 K8 tf32)\n", SMEM, tiles);
    for (int grid : {1, 76, 148}) {
        for (int mode : {0, 1, 2, 6}) {
            for (int bg = 0; bg < (mode <= 2 ? 2 : 1); ++bg) {
                CK(cudaMemset(d_cyc, 0, 148 * 8));
                CK(cudaMemset(d_bg, 0, 148 * 8));
                for (int rep = 0; rep < 2; ++rep) {
                    k_rate<<<grid, 320, SMEM>>>(mode, bg, tiles, gsrc, d_cyc, d_bg);
                    CK(cudaDeviceSynchronize());
                }
                std::vector<unsigned long long> c(grid), b(grid);
                CK(cudaMemcpy(c.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(b.data(), d_bg, grid * 8, cudaMemcpyDeviceToHost));
                double cs = 0, bs = 0, cmax = 0;
                for (int i = 0; i < grid; ++i) { cs += (double)c[i]; bs += (double)b[i]; if ((double)c[i] > cmax) cmax = (double)c[i]; }
                cs /= grid; bs /= grid;
                if (mode <= 2)
                    printf("grid %3d mode %d bg %d: %8.1f cycles/MMA (max CTA %8.1f)  bg copy %6.1f B/clk\n", grid, mode, bg, cs / (tiles * 48.0),
                           cmax / (tiles * 48.0), bs / cs);
                else
                    printf("grid %3d tcgen05.ld drain: %8.1f cycles per 3 x [128x128] accumulators (%.1f B/clk)\n", grid, cs / tiles, 3.0 * 65536.0 / (cs / tiles));
            }
        }
    }
    return 0;
}
