// tc_pipe.cuh -- the warp-specialised tcgen05 pipeline shared by the fused PaiNN node kernels (painn_fused.cu) and the pre-split-weight
// GEMM (gemm_ps.cu): weight tiles as ready-made TF32 hi / lo shared-memory images streamed by cp.async.bulk through an mbarrier ring, the
// activation operand X written by worker warps (loader functors or epilogue registers), one MMA-issuing thread (3xTF32, correction + two
// alternating main accumulators), TMEM staging of the summed accumulators, epilogues in rolled 16-column chunks.  D[feature, row] =
// W[feature, k] . X[row, k]^T: features on TMEM lanes, rows (atoms / edges / pairs) on columns.  See painn_fused.cu for the measurements.
#pragma once
#include "common.cuh"

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
constexpr int SMEM_TOTAL = SMEM_BARS + 256;
// Worker warps.  Default: all NWORK warps load operands AND run epilogues, in program order.  NF_TWO_GROUPS: warps [0, NEPI) only drain /
// run epilogues / write chained operands, warps [NEPI, NEPI + NLOAD) only run the loader functors and therefore run AHEAD of the epilogues
// (their global loads overlap epilogue work and MMAs); the groups meet at mbarriers (operand ready / free, and `dep` for data handed over
// through global memory).
#ifdef NF_TWO_GROUPS
constexpr int NEPI = NWORK / 2, NLOAD = NWORK / 2;
#else
constexpr int NEPI = NWORK, NLOAD = NWORK;
#endif
constexpr int CPT = NT / (NEPI / 4);       // accumulator columns (atoms) per epilogue thread (4 TMEM lane groups x NEPI/4 column parts)
constexpr int RPT = NT / NLOAD;            // operand rows per loader thread
constexpr int NTHREADS = 32 * (NWORK + 2); // + producer warp + MMA issuer warp
constexpr int TMEM_COLS = 4 * NT;          // three accumulators + staging
enum { U_NEWX = 1, U_FIRST = 2, U_LAST = 4, U_XLAST = 8 };
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
__device__ __forceinline__ void work_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NWORK) : "memory"); }  // all worker warps (single-group builds)
// plain (coherent) 16-byte load: for arrays written earlier in the SAME kernel (ld.global.nc / __ldg would be wrong there)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t TM_CORR = 0, TM_MAIN0 = NT, TM_MAIN1 = 2 * NT, TM_STAGE = 3 * NT;  // TMEM columns

struct Ctx {
    float *x_hi, *x_lo;
    unsigned char* ring;
    uint64_t *full, *empty, *x_ready, *x_free, *acc_full, *buf_empty, *dep;
    // xsplit: the activation operand is handed over in two K halves (k < 64: x_ready / x_free, k >= 64: x_ready2 / x_free2), so that the next
    // operand's first half is written while the MMAs still read the second half of the current one (and the MMAs start on the first half while
    // the second is written): double buffering at half-operand granularity, no extra shared memory.  r2b measurement that motivated it
    // (tools/gemm_ps_prof.py, K = 512): issuer waits X 3.5 k + workers wait X release 3.1 k of 8.9 k cycles per (N tile, K chunk).
    uint64_t *x_ready2, *x_free2;
    int xsplit = 0;
    uint32_t tmem;
    int xg = 0;  // X generations written so far (worker warps) / consumed (issuer)
    int o = 0;   // output tiles drained so far (worker warps) / committed (issuer)
    NF_PROF_DO(long long w_acc = 0, w_xfree = 0, w_x = 0, w_buf = 0, w_full = 0, t_last = 0;)
};

// producer: one thread streams the weight tiles of the units (tile_of(u) = index into the prepared buffer), STAGES_PER_TILE stages each
// (`spt` < STAGES_PER_TILE: K <= 32 * spt, the remaining stages of a tile image are zeros and are neither copied nor multiplied)
template <class TileFn>
__device__ __forceinline__ void run_producer_t(const Ctx& c, int n_units, const unsigned char* wt, TileFn tile_of, int spt = STAGES_PER_TILE) {
    int q = 0;
#pragma unroll 1
    for (int u = 0; u < n_units; ++u) {
        const unsigned char* src = wt + (size_t)tile_of(u) * WTILE_BYTES;
#pragma unroll 1
        for (int st = 0; st < spt; ++st, ++q) {
            const int slot = q % W_STAGES, use = q / W_STAGES;
            if (use > 0) mbar_wait(c.empty + slot, (uint32_t)((use - 1) & 1));
            mbar_expect_tx(c.full + slot, WST_BYTES);
            bulk_g2s(c.ring + slot * WST_BYTES, src + (size_t)st * WST_BYTES, WST_BYTES, c.full + slot);
        }
    }
}
__device__ __forceinline__ void run_producer(const Ctx& c, const Prog& prog, const unsigned char* wt) {
    run_producer_t(c, prog.n, wt, [&](int u) { return (int)prog.tile[u]; });
}

// MMA issuer: one thread walks the program.  An accumulator buffer is waited for right before its first MMA of an output tile, so the
// correction MMAs start as soon as the epilogue has read the previous tile's correction buffer.
template <class FlagFn>
__device__ __forceinline__ void run_issuer_t(Ctx& c, int n_units, FlagFn flags_of, int spt = STAGES_PER_TILE) {
    constexpr uint32_t IDESC = umma_idesc_tf32(128, NT);
    const uint64_t dx_hi0 = umma_desc(s_u32(c.x_hi), XLBO, 128), dx_lo0 = umma_desc(s_u32(c.x_lo), XLBO, 128);
    int q = 0, ks_out = 0;
#pragma unroll 1
    for (int u = 0; u < n_units; ++u) {
        const int fl = flags_of(u);
        const bool newx = (fl & U_NEWX) != 0;
        if (newx) { NF_PROF_DO(const long long t0_ = clock64();) mbar_wait(c.x_ready, (uint32_t)(c.xg & 1)); ++c.xg; NF_PROF_DO(c.w_x += clock64() - t0_;) }
        if (fl & U_FIRST) ks_out = 0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint64_t dx_hi = dx_hi0, dx_lo = dx_lo0;
#pragma unroll 1
        for (int st = 0; st < spt; ++st, ++q) {
            const int slot = q % W_STAGES;
            if (c.xsplit && newx && st == STAGES_PER_TILE / 2) {  // the second K half of a new operand
                NF_PROF_DO(const long long t0_ = clock64();)
                mbar_wait(c.x_ready2, (uint32_t)((c.xg - 1) & 1));
                NF_PROF_DO(c.w_x += clock64() - t0_;)
            }
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
            if (c.xsplit && (fl & U_XLAST) && st == STAGES_PER_TILE / 2 - 1) umma_commit(c.x_free);  // first K half: no later MMA reads it
        }
        if (fl & U_XLAST) {
            if (!c.xsplit) umma_commit(c.x_free);
            else { if (spt < STAGES_PER_TILE / 2) umma_commit(c.x_free); umma_commit(c.x_free2); }
        }
        if (fl & U_LAST) { umma_commit(c.acc_full); ++c.o; }
    }
}
__device__ __forceinline__ void run_issuer(Ctx& c, const Prog& prog) {
    run_issuer_t(c, prog.n, [&](int u) { return (int)prog.flag[u]; });
}

// worker warps: fill the activation operand with f(row 0..127 of the tile, chunk 0..31) -> 4 consecutive k values.
// Two halves of 8 rows per thread (rolled).  Per half: ALL global loads are issued (and f's arithmetic done) BEFORE the thread waits for
// the previous operand to be released, so their latency overlaps the MMAs still reading that operand; only split + 16 shared-memory
// stores follow the wait.  (8 worker warps per SM: a load -> use -> store sequence per element would expose one L2 round trip each.)
template <class Fn>
__device__ __forceinline__ void load_x(Ctx& c, int wtid, Fn f) {
    if (c.xsplit) {
        // half-operand hand-over: the two K halves must be written by DIFFERENT WARPS -- a warp whose lanes wait on two barriers reconverges
        // after the wait loop, i.e. both halves would wait for the later barrier (first version, by lane: no gain at all).  Warps [0, NLOAD/2)
        // write k < 64, the others k >= 64; a warp instruction covers 2 rows x 16 chunks (two 256-byte global segments, conflict-free
        // 16-byte shared-memory stores per quarter warp).
        const int lane = wtid & 31, wrp = wtid >> 5, half = wrp >= NLOAD / 2 ? 1 : 0, w8 = wrp - half * (NLOAD / 2);
        const int kc = (lane & 15) + 16 * half, rsub = lane >> 4;
        constexpr int ITEMS = (NT * 16) / (32 * (NLOAD / 2));  // (row, chunk) pairs per thread: 8 for NT = 128, NLOAD = 16
        static_assert(ITEMS * 32 * (NLOAD / 2) == NT * 16 && ITEMS <= 16, "load_x xsplit mapping");
        float4 t[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) t[it] = f(2 * (w8 + (NLOAD / 2) * it) + rsub, kc);
        NF_PROF_DO(const long long t0_ = clock64();)
        if (c.xg > 0) mbar_wait(half ? c.x_free2 : c.x_free, (uint32_t)((c.xg - 1) & 1));  // the MMAs that read my half of the previous operand have retired
        NF_PROF_DO(c.w_xfree += clock64() - t0_;)
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int r = 2 * (w8 + (NLOAD / 2) * it) + rsub;
            float4 hi, lo;
            split4(t[it], hi, lo);
            st4(c.x_hi + kc * XLBOF + r * 4, hi);
            st4(c.x_lo + kc * XLBOF + r * 4, lo);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(half ? c.x_ready2 : c.x_ready);
        ++c.xg;
        return;
    }
    const int kc = wtid & 31, w = wtid >> 5;
#pragma unroll 1
    for (int h = 0; h < RPT / 8; ++h) {
        float4 t[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) t[it] = f(w + NLOAD * (8 * h + it), kc);
        NF_PROF_DO(const long long t0_ = clock64();)
        if (c.xg > 0) mbar_wait(c.x_free, (uint32_t)((c.xg - 1) & 1));  // every MMA that read the previous operand has retired
        NF_PROF_DO(c.w_xfree += clock64() - t0_;)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = w + NLOAD * (8 * h + it);
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
    bool second;  // my feature row k lies in the second K half
    __device__ __forceinline__ XPut(const Ctx& c, int k) {
        second = c.xsplit && k >= 64;
        if (c.xg > 0) mbar_wait(second ? c.x_free2 : c.x_free, (uint32_t)((c.xg - 1) & 1));
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
        mbar_arrive(second ? c.x_ready2 : c.x_ready);
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
__device__ __forceinline__ Ctx setup(unsigned char* smem, int tid, int warp, int xsplit = 0) {
    Ctx c;
    c.xsplit = xsplit;
    c.x_hi = reinterpret_cast<float*>(smem);
    c.x_lo = reinterpret_cast<float*>(smem + X_BYTES);
    c.ring = smem + 2 * X_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARS);
    c.full = bars; c.empty = bars + W_STAGES; c.x_ready = bars + 2 * W_STAGES; c.x_free = c.x_ready + 1; c.acc_full = c.x_ready + 2;
    c.buf_empty = c.x_ready + 3;
    c.dep = c.x_ready + 6;  // 4 hand-over barriers between the worker groups
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.x_ready + 10);
    c.x_ready2 = c.x_ready + 11; c.x_free2 = c.x_ready + 12;
    if (tid == 0) {
        for (int s = 0; s < W_STAGES; ++s) { mbar_init(c.full + s, 1); mbar_init(c.empty + s, 1); }
        mbar_init(c.x_ready, xsplit ? 16 * NLOAD : 32 * NLOAD);  // == 32 * NEPI: one group writes a whole operand generation (half of it with xsplit)
        mbar_init(c.x_free, 1);
        mbar_init(c.x_ready2, 16 * NLOAD);
        mbar_init(c.x_free2, 1);
        mbar_init(c.acc_full, 1);
        for (int b = 0; b < 3; ++b) mbar_init(c.buf_empty + b, 32 * NEPI);
        for (int b = 0; b < 4; ++b) mbar_init(c.dep + b, 32 * NLOAD);
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

// ---- worker roles
__device__ __forceinline__ bool role_epi(int warp) { return warp < NEPI; }
#ifdef NF_TWO_GROUPS
__device__ __forceinline__ bool role_load(int warp) { return warp >= NEPI && warp < NEPI + NLOAD; }
__device__ __forceinline__ int load_tid(int tid) { return tid - 32 * NEPI; }
#else
__device__ __forceinline__ bool role_load(int warp) { return warp < NLOAD; }
__device__ __forceinline__ int load_tid(int tid) { return tid; }
#endif
// data handed from one group to the other through GLOBAL memory (k = which hand-over of the kernel, each used once): the producers arrive
// after their stores, the consumers wait; single-group builds: one CTA-wide barrier of the worker warps at the producer's point
__device__ __forceinline__ void dep_signal(const Ctx& c, int k) {
#ifdef NF_TWO_GROUPS
    mbar_arrive(c.dep + k);
#else
    (void)c; (void)k;
    work_barrier();
#endif
}
__device__ __forceinline__ void dep_wait(const Ctx& c, int k) {
#ifdef NF_TWO_GROUPS
    mbar_wait(c.dep + k, 0u);
#else
    (void)c; (void)k;
#endif
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


}  // namespace
