// painn_fused.cu -- the whole per-atom ("node") part of a PaiNN layer as ONE persistent tcgen05 kernel per direction.
//
// Replaces, per layer, the five nn.Linear GEMMs and the elementwise glue of
//   PaiNNUpdate.forward            nablaDFT/painn_pyg/painn.py:535-548   (schnetpack PaiNNMixing)
//   the message MLP x_proj         nablaDFT/painn_pyg/painn.py:459-464   (schnetpack interatomic_context_net)
//   the readout's first Linear     nablaDFT/painn_pyg/painn.py:79-83     (schnetpack Atomwise.outnet[0])
// and their autograd backward (painn.py:135-146).  Round 1 ran them as 12 separate 3xTF32 GEMM launches + 6 elementwise launches per
// layer and direction: 72 GEMM launches = 2.4 ms of a 3.4 ms step at ~0.10 of the tensor roofline, because every
// launch re-staged and re-split its activation slab, the producers split the weights again for every row slab, and the fixed costs of a
// 76-CTA launch were paid 12 times.
//
// Design (measured basis: tools/mma_rate.cu, profiles/r2_mma_rate.txt -- a tcgen05.mma.kind::tf32 M128 N128 K8 with BOTH operands in
// shared memory issues every 71 cycles next to a 57 B/clk bulk-copy stream; with the A operand in tensor memory it needs 94-106):
//   * CTA = 128 atoms.  Every GEMM of the chain is computed TRANSPOSED: D[feature, atom] = W[feature, k] . X[atom, k]^T, i.e. the weight
//     tile is the MMA's A operand (M = 128 output features) and the activations are the B operand (N = 128 atoms).  The accumulator then
//     has features on TMEM lanes and atoms on columns: an epilogue thread owns one feature and 64 atoms, so every global store / load of
//     an [atom][feature] array is a 128-byte coalesced warp access, the bias is a per-thread scalar, and writing the next activation
//     operand into shared memory ([atoms] x K, K-major) is a conflict-free 4-byte store pattern.  No transposes, no row-per-thread access.
//   * weights are split into TF32 hi / lo ONCE per call by k_prep_painn into ready-made shared-memory images (one 128 x 128 tile =
//     4 stages x [hi | lo] x 16 KB, canonical no-swizzle K-major); a producer thread streams them with one cp.async.bulk per stage
//     through a 3-stage mbarrier ring.  Nobody splits weights inside the GEMM any more.
//   * the activation operand X [128 atoms x 128 k] (hi + lo, 132 KB) is written by the 8 worker warps: either by a LOADER functor
//     (coalesced global loads, elementwise math fused in: sqrt-norm, the combine backward, silu' ...) or directly from the previous
//     GEMM's epilogue registers (silu(h) -> next operand) -- chained activations never go through global memory to be re-read as operands.
//   * one thread issues the MMAs (3xTF32: lo.hi + hi.lo into a correction accumulator, hi.hi alternating over two main accumulators so
//     that no accumulator chain is longer than 8 per 128 k: the tensor core truncates on accumulate, see gemm_tc.cu); TMEM columns
//     [0,384) = the three accumulators, [384,512) = a STAGING buffer: the epilogue first sums the three accumulators into it (compact,
//     non-inlined code), releases them to the issuer one by one, and then walks the staged tile in 16-atom chunks inside ROLLED loops.
//     (First version: v[64] per thread, everything unrolled -> 25 k SASS instructions = 400 KB executed once per CTA: the kernel was
//     instruction-fetch bound, 250 k cycles of "busy" workers.  tools/nf_prof.py, profiles/r2_fused_role_timing.md.)
//     K > 128 (backward) accumulates over several X operands in place.
//   * roles meet only through mbarriers: W ring full/empty, X ready/free, accumulator full, accumulator buffer empty.
// Forward  kernel = update(l) [+ message MLP(l+1) | readout Linear]   (15 tiles of 48 MMAs)
// Backward kernel = [message-MLP backward(l+1) | readout backward] + update backward(l)
#include "painn_node.cuh"

namespace {

constexpr int F = NB_F;
// CTA = NT atoms.  Shipped: NT = 128, one CTA per SM, 16 worker warps, 3 ring stages of 32 k.
// Tried (NF_SMALL_TILES): NT = 64 with TWO CTAs resident per SM (108 KB of shared memory, 256 TMEM columns, 320 threads each, 5 ring stages
// of 8 k) so that one CTA's store / load / TMEM-drain phases overlap its neighbour's MMAs and a 9.7 k-atom batch covers all 148 SMs:
// 1.55 ms instead of 1.35 ms per step -- an N = 64 MMA re-reads the 4 KB weight operand for half as many columns (139 cycles per MMA
// measured with two CTAs sharing the tensor pipe, 71 at N = 128), and the tensor phase became the long one.
#ifdef NF_SMALL_TILES
constexpr int NT = 64, KSTAGE = 8, W_STAGES = 5, NWORK = 8, CTAS_PER_SM = 2;
#else
constexpr int NT = 128, KSTAGE = 32, W_STAGES = 3, NWORK = 16, CTAS_PER_SM = 1;
#endif
constexpr int XLBO = NT * 16 + 16;         // bytes between 16-byte k-chunks of X (padded: the 8 chunk writers of a row hit 8 bank groups)
constexpr int XLBOF = XLBO / 4;
constexpr int X_BYTES = 32 * XLBO;         // one of hi / lo, K = 128
constexpr int WLBO = 128 * 16;             // weight stages are written by the bulk-copy engine: no padding needed
constexpr int WST_BYTES = 2 * (KSTAGE / 4) * WLBO;  // one ring stage: [hi | lo] x KSTAGE/4 chunks x 128 rows x 16 B
constexpr int STAGES_PER_TILE = 128 / KSTAGE;
constexpr int WTILE_BYTES = STAGES_PER_TILE * WST_BYTES;  // 128 rows x 128 k, hi + lo = 128 KB
constexpr int SMEM_BARS = 2 * X_BYTES + W_STAGES * WST_BYTES;
constexpr int SMEM_TOTAL = SMEM_BARS + 192;
constexpr int CPT = NT / (NWORK / 4);      // accumulator columns (atoms) per worker thread (worker warps: 4 TMEM lane groups x NWORK/4 column parts)
constexpr int RPT = NT / NWORK;            // operand rows per worker thread
constexpr int NTHREADS = 32 * (NWORK + 2); // + producer warp + MMA issuer warp
constexpr int TMEM_COLS = 4 * NT;          // three accumulators + staging
constexpr int TILES_PER_LAYER = 22;

enum { U_NEWX = 1, U_FIRST = 2, U_LAST = 4, U_XLAST = 8 };
// weight tiles of a layer (index into the prepared buffer, see k_prep_painn)
enum { T_UV = 0, T_UW, T_B1A, T_B1B, T_B2_0, T_B2_1, T_B2_2, T_A1, T_A2_0, T_A2_1, T_A2_2,
       T_B2T_0, T_B2T_1, T_B2T_2, T_B1AT, T_B1BT, T_UT_0, T_UT_1, T_A2T_0, T_A2T_1, T_A2T_2, T_A1T };

#ifdef NF_PROF
// role timing (clock64, summed over CTAs): 0 issuer total, 1 issuer waits X, 2 issuer waits TMEM buffers, 3 issuer waits W ring,
// 4 worker(thread 0) total, 5 worker waits accumulator, 6 worker waits X release, 7 CTAs       [fwd: 0..7, bwd: 8..15]
__device__ unsigned long long g_nf_prof[16];
__device__ unsigned long long g_nf_phase[64];  // worker thread 0: cycles between consecutive NF_MARK points [fwd 0..31 | bwd 32..63]
#define NF_PROF_DO(...) __VA_ARGS__
#define NF_MARK(i) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_nf_phase[NF_BASE + (i)], (unsigned long long)(now_ - c.t_last)); c.t_last = now_; } } while (0)
#else
#define NF_PROF_DO(...)
#define NF_MARK(i)
#endif

struct Prog {
    int n;
    uint16_t tile[24];
    uint8_t flag[24];
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)), "l"(src), "r"(bytes),
                 "r"(s_u32(bar))
                 : "memory");
}
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
__device__ __forceinline__ void work_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NWORK) : "memory"); }
// plain (coherent) 16-byte load: for arrays written earlier in the SAME kernel (ld.global.nc / __ldg would be wrong there)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------------------------------------------------------
// Weight preparation: every 128 x 128 block a fused kernel uses, as TF32 hi / lo shared-memory images.
// tile t of layer l lives at (l * 22 + t) * WTILE_BYTES; the two readout tiles follow the last layer.
// element (row r, k) of a tile: stage k / KSTAGE, hi first then lo (KSTAGE / 4 chunks of 2 KB each), chunk (k % KSTAGE) / 4, then r * 16 + (k % 4) * 4 bytes.
struct TileSrc { const float* p; int ld, row0, k0, trans, rows, kvalid; };

__device__ __forceinline__ TileSrc tile_src(const nb200_painn_weights& w, int idx) {
    const int L = w.n_layers;
    TileSrc s{nullptr, F, 0, 0, 0, 128, 128};
    if (idx >= L * TILES_PER_LAYER) {  // readout Linear R1 [F/2][F]: forward tile (rows = outputs, 64 valid) and transposed tile (k = outputs)
        s.p = w.R1; s.ld = F;
        if (idx - L * TILES_PER_LAYER == 0) { s.rows = F / 2; } else { s.trans = 1; s.kvalid = F / 2; }
        return s;
    }
    const int l = idx / TILES_PER_LAYER, t = idx % TILES_PER_LAYER;
    const float* A1 = w.A1 + (size_t)l * F * F;
    const float* A2 = w.A2 + (size_t)l * 3 * F * F;
    const float* U = w.U + (size_t)l * 2 * F * F;
    const float* B1 = w.B1 + (size_t)l * F * 2 * F;
    const float* B2 = w.B2 + (size_t)l * 3 * F * F;
    switch (t) {
        case T_UV: s.p = U; break;
        case T_UW: s.p = U; s.row0 = F; break;
        case T_B1A: s.p = B1; s.ld = 2 * F; break;
        case T_B1B: s.p = B1; s.ld = 2 * F; s.k0 = F; break;
        case T_B2_0: case T_B2_1: case T_B2_2: s.p = B2; s.row0 = (t - T_B2_0) * F; break;
        case T_A1: s.p = A1; break;
        case T_A2_0: case T_A2_1: case T_A2_2: s.p = A2; s.row0 = (t - T_A2_0) * F; break;
        // transposed tiles (Linear backward w.r.t. the input): element (r = input feature, k = output feature) = W[k0 + k][row0 + r]
        case T_B2T_0: case T_B2T_1: case T_B2T_2: s.p = B2; s.trans = 1; s.k0 = (t - T_B2T_0) * F; break;
        case T_B1AT: s.p = B1; s.ld = 2 * F; s.trans = 1; break;
        case T_B1BT: s.p = B1; s.ld = 2 * F; s.trans = 1; s.row0 = F; break;
        case T_UT_0: case T_UT_1: s.p = U; s.trans = 1; s.k0 = (t - T_UT_0) * F; break;
        case T_A2T_0: case T_A2T_1: case T_A2T_2: s.p = A2; s.trans = 1; s.k0 = (t - T_A2T_0) * F; break;
        default: s.p = A1; s.trans = 1; break;  // T_A1T
    }
    return s;
}

__global__ void __launch_bounds__(256) k_prep_painn(nb200_painn_weights w, unsigned char* __restrict__ dst) {
    const int idx = blockIdx.x >> 2, st = blockIdx.x & 3;
    const TileSrc s = tile_src(w, idx);
    unsigned char* tile = dst + (size_t)idx * WTILE_BYTES;  // this block: k in [32 st, 32 st + 32)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int item = threadIdx.x + 256 * i;  // 1024 (chunk, row) pairs of the stage
        int kc, r;
        if (!s.trans) { kc = item & 7; r = item >> 3; } else { r = item & 127; kc = item >> 7; }  // coalesced along the source's contiguous axis
        const int k = st * 32 + kc * 4;
        float e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = r < s.rows && k + j < s.kvalid;
            e[j] = !ok ? 0.f : !s.trans ? __ldg(s.p + (size_t)(s.row0 + r) * s.ld + s.k0 + k + j) : __ldg(s.p + (size_t)(s.k0 + k + j) * s.ld + s.row0 + r);
        }
        float4 hi, lo;
        split4(make_float4(e[0], e[1], e[2], e[3]), hi, lo);
        const int kk = 32 * st + 4 * kc;  // first k of this chunk
        float* out_hi = reinterpret_cast<float*>(tile + (size_t)(kk / KSTAGE) * WST_BYTES + (size_t)((kk % KSTAGE) / 4) * WLBO) + r * 4;
        st4(out_hi, hi);
        st4(out_hi + (KSTAGE / 4) * WLBO / 4, lo);
    }
}

// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t TM_CORR = 0, TM_MAIN0 = NT, TM_MAIN1 = 2 * NT, TM_STAGE = 3 * NT;  // TMEM columns

struct Ctx {
    float *x_hi, *x_lo;
    unsigned char* ring;
    uint64_t *full, *empty, *x_ready, *x_free, *acc_full, *buf_empty;
    uint32_t tmem;
    int xg = 0;  // X generations written so far (worker warps) / consumed (issuer)
    int o = 0;   // output tiles drained so far (worker warps) / committed (issuer)
    NF_PROF_DO(long long w_acc = 0, w_xfree = 0, w_x = 0, w_buf = 0, w_full = 0, t_last = 0;)
};

// producer: one thread streams the program's weight tiles, 4 stages each, through the ring
__device__ __forceinline__ void run_producer(const Ctx& c, const Prog& prog, const unsigned char* wt) {
    int q = 0;
    for (int u = 0; u < prog.n; ++u) {
        const unsigned char* src = wt + (size_t)prog.tile[u] * WTILE_BYTES;
#pragma unroll 1
        for (int st = 0; st < STAGES_PER_TILE; ++st, ++q) {
            const int slot = q % W_STAGES, use = q / W_STAGES;
            if (use > 0) mbar_wait(c.empty + slot, (uint32_t)((use - 1) & 1));
            mbar_expect_tx(c.full + slot, WST_BYTES);
            bulk_g2s(c.ring + slot * WST_BYTES, src + (size_t)st * WST_BYTES, WST_BYTES, c.full + slot);
        }
    }
}

// MMA issuer: one thread walks the program.  An accumulator buffer is waited for right before its first MMA of an output tile, so the
// correction MMAs start as soon as the epilogue has read the previous tile's correction buffer.
__device__ __forceinline__ void run_issuer(Ctx& c, const Prog& prog) {
    constexpr uint32_t IDESC = umma_idesc_tf32(128, NT);
    const uint64_t dx_hi0 = umma_desc(s_u32(c.x_hi), XLBO, 128), dx_lo0 = umma_desc(s_u32(c.x_lo), XLBO, 128);
    int q = 0, ks_out = 0;
#pragma unroll 1
    for (int u = 0; u < prog.n; ++u) {
        const int fl = prog.flag[u];
        if (fl & U_NEWX) { NF_PROF_DO(const long long t0_ = clock64();) mbar_wait(c.x_ready, (uint32_t)(c.xg & 1)); ++c.xg; NF_PROF_DO(c.w_x += clock64() - t0_;) }
        if (fl & U_FIRST) ks_out = 0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint64_t dx_hi = dx_hi0, dx_lo = dx_lo0;
#pragma unroll 1
        for (int st = 0; st < STAGES_PER_TILE; ++st, ++q) {
            const int slot = q % W_STAGES;
            NF_PROF_DO(const long long t1_ = clock64();)
            mbar_wait(c.full + slot, (uint32_t)((q / W_STAGES) & 1));
            NF_PROF_DO(c.w_full += clock64() - t1_;)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t wh = s_u32(c.ring + slot * WST_BYTES);
            uint64_t dw_hi = umma_desc(wh, WLBO, 128), dw_lo = umma_desc(wh + (KSTAGE / 4) * WLBO, WLBO, 128);
#pragma unroll
            for (int ks = 0; ks < KSTAGE / 8; ++ks, ++ks_out) {  // k-step of 8: lo.hi + hi.lo -> correction, hi.hi -> alternating main accumulator
                if (ks_out < 2 && c.o > 0) {  // first touch of the buffers in this output tile: the epilogue of the previous tile has read them
                    NF_PROF_DO(const long long t2_ = clock64();)
                    if (ks_out == 0) mbar_wait(c.buf_empty + 0, (uint32_t)((c.o - 1) & 1));
                    mbar_wait(c.buf_empty + 1 + ks_out, (uint32_t)((c.o - 1) & 1));
                    NF_PROF_DO(c.w_buf += clock64() - t2_;)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                umma_tf32(c.tmem + TM_CORR, dw_lo, dx_hi, IDESC, ks_out > 0 ? 1u : 0u);
                umma_tf32(c.tmem + TM_CORR, dw_hi, dx_lo, IDESC, 1u);
                umma_tf32(c.tmem + ((ks_out & 1) ? TM_MAIN1 : TM_MAIN0), dw_hi, dx_hi, IDESC, ks_out >= 2 ? 1u : 0u);
                dw_hi += (2 * WLBO) >> 4; dw_lo += (2 * WLBO) >> 4;
                dx_hi += (2 * XLBO) >> 4; dx_lo += (2 * XLBO) >> 4;
            }
            umma_commit(c.empty + slot);  // frees the ring stage when these MMAs retire
        }
        if (fl & U_XLAST) umma_commit(c.x_free);
        if (fl & U_LAST) { umma_commit(c.acc_full); ++c.o; }
    }
}

// worker warps: fill the activation operand with f(row 0..127 of the tile, chunk 0..31) -> 4 consecutive k values.
// Two halves of 8 rows per thread (rolled).  Per half: ALL global loads are issued (and f's arithmetic done) BEFORE the thread waits for
// the previous operand to be released, so their latency overlaps the MMAs still reading that operand; only split + 16 shared-memory
// stores follow the wait.  (8 worker warps per SM: a load -> use -> store sequence per element would expose one L2 round trip each.)
template <class Fn>
__device__ __forceinline__ void load_x(Ctx& c, int wtid, Fn f) {
    const int kc = wtid & 31, w = wtid >> 5;
#pragma unroll 1
    for (int h = 0; h < RPT / 8; ++h) {
        float4 t[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) t[it] = f(w + NWORK * (8 * h + it), kc);
        NF_PROF_DO(const long long t0_ = clock64();)
        if (c.xg > 0) mbar_wait(c.x_free, (uint32_t)((c.xg - 1) & 1));  // every MMA that read the previous operand has retired
        NF_PROF_DO(c.w_xfree += clock64() - t0_;)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = w + NWORK * (8 * h + it);
            float4 hi, lo;
            split4(t[it], hi, lo);
            st4(c.x_hi + kc * XLBOF + r * 4, hi);
            st4(c.x_lo + kc * XLBOF + r * 4, lo);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(c.x_ready);
    ++c.xg;
}

// worker warps, epilogue side: this thread's values (feature k, atom n) become the next operand
struct XPut {
    float *hi, *lo;
    __device__ __forceinline__ XPut(const Ctx& c, int k) {
        if (c.xg > 0) mbar_wait(c.x_free, (uint32_t)((c.xg - 1) & 1));
        hi = c.x_hi + (k >> 2) * XLBOF + (k & 3);
        lo = c.x_lo + (k >> 2) * XLBOF + (k & 3);
    }
    __device__ __forceinline__ void put(int n, float v) const {
        float h, l;
        split_tf32(v, h, l);
        hi[n * 4] = h;
        lo[n * 4] = l;
    }
    __device__ __forceinline__ void done(Ctx& c) const {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(c.x_ready);
        ++c.xg;
    }
};

#define NF_R32(R) "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]), "=r"(R[9]), "=r"(R[10]),   \
          "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]), "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]),   \
          "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]), "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]),   \
          "=r"(R[31])
#define NF_TMEM_LD32(R, TADDR)                                                                                                                \
    asm volatile(                                                                                                                             \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"   \
        "%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                                                               \
        : NF_R32(R)                                                                                                                          \
        : "r"(TADDR)                                                                                                                         \
        : "memory")
#define NF_TMEM_ST32(TADDR, R)                                                                                                                \
    asm volatile(                                                                                                                             \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"   \
        "%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(TADDR),                                                                                    \
        "r"(R[0]), "r"(R[1]), "r"(R[2]), "r"(R[3]), "r"(R[4]), "r"(R[5]), "r"(R[6]), "r"(R[7]), "r"(R[8]), "r"(R[9]), "r"(R[10]), "r"(R[11]),        \
        "r"(R[12]), "r"(R[13]), "r"(R[14]), "r"(R[15]), "r"(R[16]), "r"(R[17]), "r"(R[18]), "r"(R[19]), "r"(R[20]), "r"(R[21]), "r"(R[22]),         \
        "r"(R[23]), "r"(R[24]), "r"(R[25]), "r"(R[26]), "r"(R[27]), "r"(R[28]), "r"(R[29]), "r"(R[30]), "r"(R[31])                                  \
        : "memory")

// worker warps: wait for output tile `o`, RN-sum its three accumulators (correction + two main) into the staging columns of this thread's
// TMEM lane, releasing each accumulator to the issuer as soon as it is in registers.  One copy of this code for all 15 call sites.
__device__ __noinline__ void drain_to_stage(uint32_t tmem, uint64_t* acc_full, uint64_t* buf_empty, int o, int warp, int add_stage) {
    mbar_wait(acc_full, (uint32_t)(o & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * CPT);
#pragma unroll 1
    for (int h = 0; h < CPT / 32; ++h) {  // passes of 32 of this thread's columns: 32 + 32 live registers
        uint32_t acc[32], r[32];
        NF_TMEM_LD32(acc, base + TM_CORR + h * 32);
        NF_TMEM_LD32(r, base + TM_MAIN0 + h * 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(r[i]));
        NF_TMEM_LD32(r, base + TM_MAIN1 + h * 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(r[i]));
        if (add_stage) {  // K > 128 split over two output tiles (forward g1pre): the first half waits in the staging columns
            NF_TMEM_LD32(r, base + TM_STAGE + h * 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(r[i]));
        }
        if (h == CPT / 32 - 1) {  // all columns of the three buffers are in registers / staged: hand them back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(buf_empty + 0); mbar_arrive(buf_empty + 1); mbar_arrive(buf_empty + 2);
        }
        NF_TMEM_ST32(base + TM_STAGE + h * 32, acc);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void drain(Ctx& c, int warp, int add_stage = 0) {
    NF_PROF_DO(const long long t0_ = clock64();)
    drain_to_stage(c.tmem, c.acc_full, c.buf_empty, c.o, warp, add_stage);
    NF_PROF_DO(c.w_acc += clock64() - t0_;)
    ++c.o;
}

// 16 staged values of this thread: atoms CPT (warp >> 2) + 16 cb .. + 15 of its feature
__device__ __forceinline__ void stage_ld16(const Ctx& c, int warp, int cb, float (&v)[16]) {
    uint32_t r[16];
    const uint32_t taddr = c.tmem + ((uint32_t)((warp & 3) * 32) << 16) + TM_STAGE + (uint32_t)((warp >> 2) * CPT + cb * 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void prog_add(Prog& p, int tile, int flags) {
    p.tile[p.n] = (uint16_t)tile;
    p.flag[p.n] = (uint8_t)flags;
    ++p.n;
}

// common prologue: carve shared memory, init barriers, allocate TMEM
__device__ __forceinline__ Ctx setup(unsigned char* smem, int tid, int warp) {
    Ctx c;
    c.x_hi = reinterpret_cast<float*>(smem);
    c.x_lo = reinterpret_cast<float*>(smem + X_BYTES);
    c.ring = smem + 2 * X_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARS);
    c.full = bars; c.empty = bars + W_STAGES; c.x_ready = bars + 2 * W_STAGES; c.x_free = c.x_ready + 1; c.acc_full = c.x_ready + 2;
    c.buf_empty = c.x_ready + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.x_ready + 6);
    if (tid == 0) {
        for (int s = 0; s < W_STAGES; ++s) { mbar_init(c.full + s, 1); mbar_init(c.empty + s, 1); }
        mbar_init(c.x_ready, 32 * NWORK);
        mbar_init(c.x_free, 1);
        mbar_init(c.acc_full, 1);
        for (int b = 0; b < 3; ++b) mbar_init(c.buf_empty + b, 32 * NWORK);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    c.tmem = *tmem_slot;
    return c;
}
__device__ __forceinline__ void teardown(const Ctx& c, int warp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "n"(TMEM_COLS) : "memory");
}

// epilogue loop over this thread's part of the staged tile: chunks of 16 atoms, rolled (one copy of the body in the instruction cache)
template <class Body>
__device__ __forceinline__ void epi_chunks(const Ctx& c, int warp, Body body) {
#pragma unroll 1
    for (int cb = 0; cb < CPT / 16; ++cb) {
        float v[16];
        stage_ld16(c, warp, cb, v);
        body(cb, v);
    }
}

// ================================================================================================== forward
struct FwdParams {
    int n_atoms, do_upd, do_mlp, do_ro;
    const unsigned char* wt;  // prepared weight tiles
    int tile_upd, tile_mlp, tile_ro;  // first tile of the layer updated / of the layer whose message MLP runs / readout forward tile
    // update (layer l): inputs after the message kernel, saved activations, outputs
    const float *q_mid, *mu_mid, *d1, *d2;
    float *VW, *nrm, *dot, *g1pre, *y, *q_next, *mu_next;  // dot = <V, Wv> per (atom, channel): saved for the backward
    float eps;
    // message MLP (layer l + 1; layer 0 when !do_upd): input when it is not produced in-kernel, saved pre-activation, output
    const float *q_mlp_in, *c1;
    float *h1pre, *xh;
    float* ro_pre;  // readout: [N, F/2] WITHOUT the bias e1 (k_readout adds it)
};

__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM) k_node_fwd(const FwdParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ Prog prog;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        prog.n = 0;
        if (P.do_upd) {
            for (int x = 0; x < 3; ++x) {
                prog_add(prog, P.tile_upd + T_UV, U_NEWX | U_FIRST | U_LAST);
                prog_add(prog, P.tile_upd + T_UW, U_FIRST | U_LAST | U_XLAST);
            }
            prog_add(prog, P.tile_upd + T_B1A, U_NEWX | U_FIRST | U_LAST | U_XLAST);
            prog_add(prog, P.tile_upd + T_B1B, U_NEWX | U_FIRST | U_LAST | U_XLAST);
            prog_add(prog, P.tile_upd + T_B2_1, U_NEWX | U_FIRST | U_LAST);
            prog_add(prog, P.tile_upd + T_B2_0, U_FIRST | U_LAST);
            prog_add(prog, P.tile_upd + T_B2_2, U_FIRST | U_LAST | U_XLAST);
        }
        if (P.do_mlp) {
            prog_add(prog, P.tile_mlp + T_A1, U_NEWX | U_FIRST | U_LAST | U_XLAST);
            prog_add(prog, P.tile_mlp + T_A2_0, U_NEWX | U_FIRST | U_LAST);
            prog_add(prog, P.tile_mlp + T_A2_1, U_FIRST | U_LAST);
            prog_add(prog, P.tile_mlp + T_A2_2, U_FIRST | U_LAST | U_XLAST);
        }
        if (P.do_ro) prog_add(prog, P.tile_ro, U_NEWX | U_FIRST | U_LAST | U_XLAST);
    }
    Ctx c = setup(smem, tid, warp);
    NF_PROF_DO(const long long tk0_ = clock64(); c.t_last = tk0_;)
#define NF_BASE 0

    if (warp == NWORK) {
        if (lane == 0) run_producer(c, prog, P.wt);
    } else if (warp == NWORK + 1) {
        if (lane == 0) {
            run_issuer(c, prog);
            NF_PROF_DO(atomicAdd(&g_nf_prof[0], (unsigned long long)(clock64() - tk0_)); atomicAdd(&g_nf_prof[1], (unsigned long long)c.w_x);
                       atomicAdd(&g_nf_prof[2], (unsigned long long)c.w_buf); atomicAdd(&g_nf_prof[3], (unsigned long long)c.w_full);)
        }
    } else {
        const int N = P.n_atoms, A0 = blockIdx.x * NT;
        const int fl = 32 * (warp & 3) + lane;   // my feature inside a 128-row weight tile
        const int n0 = CPT * (warp >> 2);        // my first atom column
        if (P.do_upd) {
            // ---- VW[(atom, x)] = mu_mid[(atom, x)] . U^T : V half, W half per cartesian component
#pragma unroll 1
            for (int x = 0; x < 3; ++x) {
                load_x(c, tid, [&](int r, int kc) { return A0 + r < N ? ldg4(P.mu_mid + (size_t)(A0 + r) * (3 * F) + x * F + 4 * kc) : f4(0.f); });
                NF_MARK(0);
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    drain(c, warp);
                    NF_MARK(1);
                    epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                        float* dst = P.VW + (size_t)(A0 + n0 + 16 * cb) * (6 * F) + x * 2 * F + half * F + fl;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (A0 + n0 + 16 * cb + j < N) dst[(size_t)j * (6 * F)] = v[j];
                    });
                    NF_MARK(21);
                }
            }
            NF_MARK(2);
            work_barrier();  // VW of this tile is visible to the loader-mapped threads below
            // ---- g1pre = [q_mid | nrm] . B1^T + d1 as two K = 128 halves summed in the staging columns
            load_x(c, tid, [&](int r, int kc) { return A0 + r < N ? ldg4(P.q_mid + (size_t)(A0 + r) * F + 4 * kc) : f4(0.f); });
            NF_MARK(3);
            {   // while the tensor core works on q_mid: nrm = sqrt(sum_x V_x^2 + eps) and dot = sum_x V_x Wv_x, 2 atoms per round
                // (keeping |V|^2 and <V,Wv> in registers across the U tiles was tried: 64 persistent registers spill, and with 230 KB of
                //  shared memory there is no L1 left for local memory -- 1.42 ms instead of 1.37 ms per step for the node kernels)
                const int kc = tid & 31, w = tid >> 5;
#pragma unroll 1
                for (int it0 = 0; it0 < RPT; it0 += 2) {
                    float4 V[2][3], Wv[2][3];
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int a = A0 + w + NWORK * (it0 + b);
                        const float* vv = P.VW + (size_t)min(a, N - 1) * (6 * F) + 4 * kc;
#pragma unroll
                        for (int x = 0; x < 3; ++x) { V[b][x] = ld4(vv + x * 2 * F); Wv[b][x] = ld4(vv + x * 2 * F + F); }
                    }
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int a = A0 + w + NWORK * (it0 + b);
                        if (a < N) {
                            float4 sq = V[b][0] * V[b][0]; fma4(sq, V[b][1], V[b][1]); fma4(sq, V[b][2], V[b][2]);
                            float4 dt = f4(0.f); fma4(dt, V[b][0], Wv[b][0]); fma4(dt, V[b][1], Wv[b][1]); fma4(dt, V[b][2], Wv[b][2]);
                            st4(P.nrm + (size_t)a * F + 4 * kc, make_float4(sqrtf(sq.x + P.eps), sqrtf(sq.y + P.eps), sqrtf(sq.z + P.eps), sqrtf(sq.w + P.eps)));
                            st4(P.dot + (size_t)a * F + 4 * kc, dt);
                        }
                    }
                }
            }
            NF_MARK(4);
            work_barrier();  // nrm (read back by the same threads) and dot (read by the y2 epilogue threads) are visible
            load_x(c, tid, [&](int r, int kc) { return A0 + r < N ? ld4(P.nrm + (size_t)(A0 + r) * F + 4 * kc) : f4(0.f); });
            NF_MARK(5);
            drain(c, warp);      // q_mid half: stays in the staging columns
            NF_MARK(6);
            NF_MARK(7);
            {
                drain(c, warp, 1);   // + nrm half
                NF_MARK(8);
                const float b = __ldg(P.d1 + fl);
                const XPut xp(c, fl);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {   // operand first: the tensor core restarts before anything is stored
#pragma unroll
                    for (int j = 0; j < 16; ++j) xp.put(n0 + 16 * cb + j, A0 + n0 + 16 * cb + j < N ? siluf_(v[j] + b) : 0.f);
                });
                xp.done(c);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float* g = P.g1pre + (size_t)(A0 + n0 + 16 * cb) * F + fl;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (A0 + n0 + 16 * cb + j < N) g[(size_t)j * F] = v[j] + b;
                });
            }
            NF_MARK(9);
            // ---- y = silu(g1pre) . B2^T + d2, tiles in the order (gate y1, scalar y0, dot-scale y2)
            {   // y1: mu_next = mu_mid + y1 * Wv   (runs while the tensor core works on the y0 / y2 tiles)
                drain(c, warp);
                NF_MARK(10);
                const float b = __ldg(P.d2 + F + fl);
#pragma unroll 1
                for (int cb = 0; cb < CPT / 8; ++cb) {  // 8 atoms per round: 48 loads in flight per thread
                    float v16[16];
                    stage_ld16(c, warp, cb >> 1, v16);
                    float tw[8][3], tm[8][3];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int a = min(A0 + n0 + 8 * cb + jj, N - 1);
                        const float* vw = P.VW + (size_t)a * (6 * F) + F + fl;
                        const float* mm = P.mu_mid + (size_t)a * (3 * F) + fl;
#pragma unroll
                        for (int x = 0; x < 3; ++x) { tw[jj][x] = vw[x * 2 * F]; tm[jj][x] = __ldg(mm + x * F); }
                    }
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int a = A0 + n0 + 8 * cb + jj;
                        if (a < N) {
                            const float y1 = ((cb & 1) ? v16[8 + jj] : v16[jj]) + b;
                            P.y[(size_t)a * (3 * F) + F + fl] = y1;
                            float* mo = P.mu_next + (size_t)a * (3 * F) + fl;
#pragma unroll
                            for (int x = 0; x < 3; ++x) mo[x * F] = fmaf(y1, tw[jj][x], tm[jj][x]);
                        }
                    }
                }
            }
            NF_MARK(11);
            {   // y0: stored, and q_next <- q_mid + y0 (completed by the y2 tile)
                drain(c, warp);
                NF_MARK(12);
                const float b = __ldg(P.d2 + fl);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = __ldg(P.q_mid + (size_t)min(A0 + n0 + 16 * cb + j, N - 1) * F + fl);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int a = A0 + n0 + 16 * cb + j;
                        if (a < N) {
                            const float y0 = v[j] + b;
                            P.y[(size_t)a * (3 * F) + fl] = y0;
                            P.q_next[(size_t)a * F + fl] = t[j] + y0;
                        }
                    }
                });
            }
            NF_MARK(13);
            {   // y2: q_next = (q_mid + y0) + y2 * <V, Wv>; it is the next operand (message MLP of the next layer / readout)
                drain(c, warp);
                NF_MARK(14);
                const float b = __ldg(P.d2 + 2 * F + fl);
                const XPut xp(c, fl);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float tq[16], td[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const size_t a = (size_t)min(A0 + n0 + 16 * cb + j, N - 1);
                        tq[j] = P.q_next[a * F + fl];
                        td[j] = P.dot[a * F + fl];
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int a = A0 + n0 + 16 * cb + j;
                        float qn = 0.f;
                        if (a < N) {
                            qn = fmaf(v[j] + b, td[j], tq[j]);
                            P.q_next[(size_t)a * F + fl] = qn;
                        }
                        xp.put(n0 + 16 * cb + j, qn);
                    }
                });
                xp.done(c);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (A0 + n0 + 16 * cb + j < N) P.y[(size_t)(A0 + n0 + 16 * cb + j) * (3 * F) + 2 * F + fl] = v[j] + b;
                });
            }
        } else {
            load_x(c, tid, [&](int r, int kc) { return A0 + r < N ? ldg4(P.q_mlp_in + (size_t)(A0 + r) * F + 4 * kc) : f4(0.f); });
        }
        NF_MARK(15);
        if (P.do_mlp) {
            {   // h1pre = q . A1^T + c1 ; silu -> operand (first), then the saved pre-activation
                drain(c, warp);
                NF_MARK(16);
                const float b = __ldg(P.c1 + fl);
                const XPut xp(c, fl);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) xp.put(n0 + 16 * cb + j, siluf_(v[j] + b));
                });
                xp.done(c);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (A0 + n0 + 16 * cb + j < N) P.h1pre[(size_t)(A0 + n0 + 16 * cb + j) * F + fl] = v[j] + b;
                });
            }
            NF_MARK(17);
#pragma unroll 1
            for (int ct = 0; ct < 3; ++ct) {  // xh = act . A2^T  (bias c2 is added inside the message kernel)
                drain(c, warp);
                NF_MARK(18);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float* dst = P.xh + (size_t)(A0 + n0 + 16 * cb) * (3 * F) + ct * F + fl;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (A0 + n0 + 16 * cb + j < N) dst[(size_t)j * (3 * F)] = v[j];
                });
            }
        }
        NF_MARK(19);
        if (P.do_ro) {
            drain(c, warp);
            epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                if (fl < F / 2) {
                    float* dst = P.ro_pre + (size_t)(A0 + n0 + 16 * cb) * (F / 2) + fl;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (A0 + n0 + 16 * cb + j < N) dst[(size_t)j * (F / 2)] = v[j];
                }
            });
        }
    }
    NF_MARK(20);
#undef NF_BASE
    NF_PROF_DO(if (tid == 0) { atomicAdd(&g_nf_prof[4], (unsigned long long)(clock64() - tk0_)); atomicAdd(&g_nf_prof[5], (unsigned long long)c.w_acc);
                            atomicAdd(&g_nf_prof[6], (unsigned long long)c.w_xfree); atomicAdd(&g_nf_prof[7], 1ull); })
    teardown(c, warp);
}

// ================================================================================================== backward
struct BwdParams {
    int n_atoms, do_mlp, do_ro, do_upd;
    const unsigned char* wt;
    int tile_mlp, tile_ro, tile_upd;  // layer whose message MLP is differentiated / readout transposed tile / layer whose update is differentiated
    // gradients: gq_a = dE/dq in (from the previous backward step) and out (dE/dq_mid of the updated layer); gq_b = scratch (dE/dq_in of
    // the layer above); cur = dE/dmu, in / out; g_xh = dE/dxh written by the message backward of the layer above
    float *gq_a, *gq_b, *cur, *gn, *gdot;  // gn holds s = gn / nrm, gdot = gq_b * y2 (scratch of this kernel)
    const float *g_xh, *h1pre, *dot;
    const float *ro_pre, *R2;      // readout backward: g_ro = R2 * silu'(ro_pre)   (ro_pre holds the biased pre-activation)
    const float *y, *VW, *nrm, *g1pre;
};

__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM) k_node_bwd(const BwdParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ Prog prog;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        prog.n = 0;
        if (P.do_mlp) {
            prog_add(prog, P.tile_mlp + T_A2T_0, U_NEWX | U_FIRST | U_XLAST);
            prog_add(prog, P.tile_mlp + T_A2T_1, U_NEWX | U_XLAST);
            prog_add(prog, P.tile_mlp + T_A2T_2, U_NEWX | U_XLAST | U_LAST);
            prog_add(prog, P.tile_mlp + T_A1T, U_NEWX | U_FIRST | U_LAST | U_XLAST);
        } else if (P.do_ro) {
            prog_add(prog, P.tile_ro, U_NEWX | U_FIRST | U_LAST | U_XLAST);
        }
        if (P.do_upd) {
            prog_add(prog, P.tile_upd + T_B2T_0, U_NEWX | U_FIRST | U_XLAST);
            prog_add(prog, P.tile_upd + T_B2T_1, U_NEWX | U_XLAST);
            prog_add(prog, P.tile_upd + T_B2T_2, U_NEWX | U_XLAST | U_LAST);
            prog_add(prog, P.tile_upd + T_B1AT, U_NEWX | U_FIRST | U_LAST);
            prog_add(prog, P.tile_upd + T_B1BT, U_FIRST | U_LAST | U_XLAST);
            for (int x = 0; x < 3; ++x) {
                prog_add(prog, P.tile_upd + T_UT_0, U_NEWX | U_FIRST | U_XLAST);
                prog_add(prog, P.tile_upd + T_UT_1, U_NEWX | U_XLAST | U_LAST);
            }
        }
    }
    Ctx c = setup(smem, tid, warp);
    NF_PROF_DO(const long long tk0_ = clock64();)

    if (warp == NWORK) {
        if (lane == 0) run_producer(c, prog, P.wt);
    } else if (warp == NWORK + 1) {
        if (lane == 0) {
            run_issuer(c, prog);
            NF_PROF_DO(atomicAdd(&g_nf_prof[8], (unsigned long long)(clock64() - tk0_)); atomicAdd(&g_nf_prof[9], (unsigned long long)c.w_x);
                       atomicAdd(&g_nf_prof[10], (unsigned long long)c.w_buf); atomicAdd(&g_nf_prof[11], (unsigned long long)c.w_full);)
        }
    } else {
        const int N = P.n_atoms, A0 = blockIdx.x * NT;
        const int fl = 32 * (warp & 3) + lane;
        const int n0 = CPT * (warp >> 2);
        if (P.do_mlp) {
            // ---- gt = g_xh . A2 (K = 384) ; gt *= silu'(h1pre) ; gq_b = gq_a + gt . A1
#pragma unroll 1
            for (int ck = 0; ck < 3; ++ck)
                load_x(c, tid, [&](int r, int kc) { return A0 + r < N ? ldg4(P.g_xh + (size_t)(A0 + r) * (3 * F) + ck * F + 4 * kc) : f4(0.f); });
            {
                drain(c, warp);
                const XPut xp(c, fl);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = __ldg(P.h1pre + (size_t)min(A0 + n0 + 16 * cb + j, N - 1) * F + fl);
#pragma unroll
                    for (int j = 0; j < 16; ++j) xp.put(n0 + 16 * cb + j, A0 + n0 + 16 * cb + j < N ? v[j] * dsiluf_(t[j]) : 0.f);
                });
                xp.done(c);
            }
        } else if (P.do_ro) {
            // ---- gq_b = g_ro . R1 with g_ro[k] = R2[k] silu'(ro_pre[k]), k < F/2 (zero-padded to K = 128)
            load_x(c, tid, [&](int r, int kc) {
                if (A0 + r >= N || kc >= F / 8) return f4(0.f);
                const float4 p = ldg4(P.ro_pre + (size_t)(A0 + r) * (F / 2) + 4 * kc), w2 = ldg4(P.R2 + 4 * kc);
                return make_float4(w2.x * dsiluf_(p.x), w2.y * dsiluf_(p.y), w2.z * dsiluf_(p.z), w2.w * dsiluf_(p.w));
            });
        }
        if (P.do_mlp || P.do_ro) {  // gq_b = dE/dq_in of the layer above; gdot = gq_b * y2 is what the combine backward needs three times
            drain(c, warp);
            epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                float t[16], ty[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const size_t a = (size_t)min(A0 + n0 + 16 * cb + j, N - 1);
                    t[j] = P.do_mlp ? P.gq_a[a * F + fl] : 0.f;
                    ty[j] = P.do_upd ? __ldg(P.y + a * (3 * F) + 2 * F + fl) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int a = A0 + n0 + 16 * cb + j;
                    if (a < N) {
                        const float g = t[j] + v[j];
                        P.gq_b[(size_t)a * F + fl] = g;
                        if (P.do_upd) P.gdot[(size_t)a * F + fl] = g * ty[j];
                    }
                }
            });
        }
        if (P.do_upd) {
            work_barrier();  // gq_b, gdot visible to the loader-mapped threads
            // ---- gt = gy . B2 (K = 384) with gy = (gq, sum_x cur_x Wv_x, gq <V, Wv>) formed on the fly (combine backward)
            load_x(c, tid, [&](int r, int kc) { return A0 + r < N ? ld4(P.gq_b + (size_t)(A0 + r) * F + 4 * kc) : f4(0.f); });
            load_x(c, tid, [&](int r, int kc) {
                if (A0 + r >= N) return f4(0.f);
                const float* vw = P.VW + (size_t)(A0 + r) * (6 * F) + F + 4 * kc;
                const float* gm = P.cur + (size_t)(A0 + r) * (3 * F) + 4 * kc;
                float4 sacc = f4(0.f);
#pragma unroll
                for (int x = 0; x < 3; ++x) fma4(sacc, ld4(gm + x * F), ldg4(vw + x * 2 * F));
                return sacc;
            });
            load_x(c, tid, [&](int r, int kc) {
                return A0 + r < N ? ld4(P.gq_b + (size_t)(A0 + r) * F + 4 * kc) * ldg4(P.dot + (size_t)(A0 + r) * F + 4 * kc) : f4(0.f);
            });
            {
                drain(c, warp);
                const XPut xp(c, fl);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = __ldg(P.g1pre + (size_t)min(A0 + n0 + 16 * cb + j, N - 1) * F + fl);
#pragma unroll
                    for (int j = 0; j < 16; ++j) xp.put(n0 + 16 * cb + j, A0 + n0 + 16 * cb + j < N ? v[j] * dsiluf_(t[j]) : 0.f);
                });
                xp.done(c);
            }
            {   // gq_a = gq_b + gt . B1[:, :F]   (dE/dq_mid of this layer: what the message backward reads)
                drain(c, warp);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = P.gq_b[(size_t)min(A0 + n0 + 16 * cb + j, N - 1) * F + fl];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (A0 + n0 + 16 * cb + j < N) P.gq_a[(size_t)(A0 + n0 + 16 * cb + j) * F + fl] = t[j] + v[j];
                });
            }
            {   // gn = gt . B1[:, F:], stored as s = gn / nrm (norm backward: gV_x += s V_x)
                drain(c, warp);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = __ldg(P.nrm + (size_t)min(A0 + n0 + 16 * cb + j, N - 1) * F + fl);
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (A0 + n0 + 16 * cb + j < N) P.gn[(size_t)(A0 + n0 + 16 * cb + j) * F + fl] = v[j] / t[j];
                });
            }
            work_barrier();  // s visible
            // ---- cur_x += gVW_x . U (K = 256: V chunk then Wv chunk), gVW formed on the fly (combine + norm backward)
#pragma unroll 1
            for (int x = 0; x < 3; ++x) {
                load_x(c, tid, [&](int r, int kc) {  // gV = gdot * Wv + s * V
                    if (A0 + r >= N) return f4(0.f);
                    const size_t a = (size_t)(A0 + r);
                    float4 o = ld4(P.gdot + a * F + 4 * kc) * ldg4(P.VW + a * (6 * F) + x * 2 * F + F + 4 * kc);
                    fma4(o, ld4(P.gn + a * F + 4 * kc), ldg4(P.VW + a * (6 * F) + x * 2 * F + 4 * kc));
                    return o;
                });
                load_x(c, tid, [&](int r, int kc) {  // gWv = cur_x * y1 + gdot * V
                    if (A0 + r >= N) return f4(0.f);
                    const size_t a = (size_t)(A0 + r);
                    float4 o = ld4(P.cur + a * (3 * F) + x * F + 4 * kc) * ldg4(P.y + a * (3 * F) + F + 4 * kc);
                    fma4(o, ld4(P.gdot + a * F + 4 * kc), ldg4(P.VW + a * (6 * F) + x * 2 * F + 4 * kc));
                    return o;
                });
                drain(c, warp);
                epi_chunks(c, warp, [&](int cb, float (&v)[16]) {
                    float t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = P.cur[(size_t)min(A0 + n0 + 16 * cb + j, N - 1) * (3 * F) + x * F + fl];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (A0 + n0 + 16 * cb + j < N) P.cur[(size_t)(A0 + n0 + 16 * cb + j) * (3 * F) + x * F + fl] = t[j] + v[j];
                });
            }
        }
    }
    NF_PROF_DO(if (tid == 0) { atomicAdd(&g_nf_prof[12], (unsigned long long)(clock64() - tk0_)); atomicAdd(&g_nf_prof[13], (unsigned long long)c.w_acc);
                            atomicAdd(&g_nf_prof[14], (unsigned long long)c.w_xfree); atomicAdd(&g_nf_prof[15], 1ull); })
    teardown(c, warp);
}

template <class K>
int set_smem(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) == cudaSuccess ? NB200_OK : nb_check_launch();
}

}  // namespace

#ifdef NF_PROF
extern "C" int nb200_debug_nf_prof(unsigned long long* out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, g_nf_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_nf_prof, z, sizeof(z)); }
    return 0;
}
extern "C" int nb200_debug_nf_phase(unsigned long long* out64, int reset) {
    if (cudaMemcpyFromSymbol(out64, g_nf_phase, sizeof(unsigned long long) * 64) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_nf_phase, z, sizeof(z)); }
    return 0;
}
#endif

int64_t nb_fused_wtile_bytes(int n_layers) { return (int64_t)(n_layers * TILES_PER_LAYER + 2) * WTILE_BYTES; }

int nb_fused_prep(const nb200_painn_weights* w, void* wtiles, cudaStream_t s) {
    const int n_tiles = w->n_layers * TILES_PER_LAYER + 2;
    k_prep_painn<<<n_tiles * 4, 256, 0, s>>>(*w, static_cast<unsigned char*>(wtiles));
    return nb_check_launch();
}

int nb_fused_node_fwd(const NbFusedFwd& a, cudaStream_t s) {
    static bool attr = false;
    if (!attr) { if (set_smem(k_node_fwd) != NB200_OK) return NB200_ECUDA; attr = true; }
    FwdParams P{};
    P.n_atoms = a.n_atoms; P.do_upd = a.layer_upd >= 0; P.do_mlp = a.layer_mlp >= 0; P.do_ro = a.readout;
    P.wt = static_cast<const unsigned char*>(a.wtiles);
    P.tile_upd = a.layer_upd * TILES_PER_LAYER; P.tile_mlp = a.layer_mlp * TILES_PER_LAYER; P.tile_ro = a.n_layers * TILES_PER_LAYER;
    P.q_mid = a.q_mid; P.mu_mid = a.mu_mid; P.d1 = a.d1; P.d2 = a.d2; P.VW = a.VW; P.nrm = a.nrm; P.dot = a.dot; P.g1pre = a.g1pre; P.y = a.y;
    P.q_next = a.q_next; P.mu_next = a.mu_next; P.eps = a.eps; P.q_mlp_in = a.q_mlp_in; P.c1 = a.c1; P.h1pre = a.h1pre; P.xh = a.xh;
    P.ro_pre = a.ro_pre;
    if (a.n_atoms <= 0) return NB200_OK;
    k_node_fwd<<<(a.n_atoms + NT - 1) / NT, NTHREADS, SMEM_TOTAL, s>>>(P);
    return nb_check_launch();
}

int nb_fused_node_bwd(const NbFusedBwd& a, cudaStream_t s) {
    static bool attr = false;
    if (!attr) { if (set_smem(k_node_bwd) != NB200_OK) return NB200_ECUDA; attr = true; }
    BwdParams P{};
    P.n_atoms = a.n_atoms; P.do_mlp = a.layer_mlp >= 0; P.do_ro = a.readout; P.do_upd = a.layer_upd >= 0;
    P.wt = static_cast<const unsigned char*>(a.wtiles);
    P.tile_mlp = a.layer_mlp * TILES_PER_LAYER; P.tile_ro = a.n_layers * TILES_PER_LAYER + 1; P.tile_upd = a.layer_upd * TILES_PER_LAYER;
    P.gq_a = a.gq_a; P.gq_b = a.gq_b; P.cur = a.cur; P.gn = a.gn; P.gdot = a.gdot; P.dot = a.dot; P.g_xh = a.g_xh; P.h1pre = a.h1pre; P.ro_pre = a.ro_pre; P.R2 = a.R2;
    P.y = a.y; P.VW = a.VW; P.nrm = a.nrm; P.g1pre = a.g1pre;
    if (a.n_atoms <= 0) return NB200_OK;
    k_node_bwd<<<(a.n_atoms + NT - 1) / NT, NTHREADS, SMEM_TOTAL, s>>>(P);
    return nb_check_launch();
}
