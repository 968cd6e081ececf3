// painn_msg.cu -- PaiNN message + segmented scatter (K_msg), forward and analytic backward.
//
// Replaces  schnetpack PaiNNInteraction.forward (config/model/painn.yaml) and
//           PaiNNMessage.forward/message/aggregate (nablaDFT/painn_pyg/painn.py:475-509):
//   gather xh[j], mu[j] -> multiply by the per-edge filter -> two scatter-adds (atomics) ->
//   residual add, with [E,384] / [E,3,128] temporaries in HBM.
// Here: one warp per atom, lane = 4 channels (float4), CSR rows streamed once, the sums are
// carried in registers and written once -- deterministic, no atomics, no temporaries.
//
// Canonical chunk roles (host permutes PaiNN-OC weights into them):
//   (a, b, c) = split(xh_j + bias),  (Wa, Wb, Wc) = split(W_e)
//   dq_i  = sum_e Wa*a ;  dmu_i[x] = sum_e (Wb*b) u_e[x] + (Wc*c) * mu_j[x]
//
// Algorithmic HBM bytes (definition A), F = 128, fp32:
//   forward : N*10F*4 + E*(3F*4 + 20)           = 5120 N + 1556 E
//   backward: N*16F*4 + E*(6F*4 + 32)           = 8192 N + 3104 E
#include "common.cuh"

#define MSG_WARPS 8
#define MSG_THREADS (MSG_WARPS * 32)

__global__ void __launch_bounds__(MSG_THREADS) k_painn_msg_fwd(const float* __restrict__ xh, const float* __restrict__ xh_bias,
                                                              const float* q, const float* __restrict__ mu,
                                                              const float* __restrict__ W, const float* __restrict__ geom,
                                                              const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                              int n_atoms, float* q_out, float* __restrict__ mu_out) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * MSG_WARPS + (threadIdx.x >> 5);
    if (i >= n_atoms) return;
    const int c = lane * 4;
    const float4 ba = ldg4(xh_bias + c), bb = ldg4(xh_bias + NB_F + c), bc = ldg4(xh_bias + 2 * NB_F + c);
    float4 dq = f4(0.f), dm0 = f4(0.f), dm1 = f4(0.f), dm2 = f4(0.f);
    const int e0 = row_ptr[i], e1 = row_ptr[i + 1];
#pragma unroll 2
    for (int e = e0; e < e1; ++e) {
        const int j = __ldg(col + e);
        const float4 g = ldg4(geom + 4 * (size_t)e);
        const float* we = W + (size_t)e * (3 * NB_F) + c;
        const float4 wa = ldg4_stream(we), wb = ldg4_stream(we + NB_F), wc = ldg4_stream(we + 2 * NB_F);
        const float* xj = xh + (size_t)j * (3 * NB_F) + c;
        const float4 a = ldg4(xj) + ba, b = ldg4(xj + NB_F) + bb, cc = ldg4(xj + 2 * NB_F) + bc;
        const float* mj = mu + (size_t)j * (3 * NB_F) + c;
        const float4 m0 = ldg4(mj), m1 = ldg4(mj + NB_F), m2 = ldg4(mj + 2 * NB_F);
        fma4(dq, wa, a);
        const float4 pb = wb * b, pc = wc * cc;
        fma4s(dm0, pb, g.x); fma4(dm0, pc, m0);
        fma4s(dm1, pb, g.y); fma4(dm1, pc, m1);
        fma4s(dm2, pb, g.z); fma4(dm2, pc, m2);
    }
    const size_t qi = (size_t)i * NB_F + c, mi = (size_t)i * (3 * NB_F) + c;
    st4(q_out + qi, *reinterpret_cast<const float4*>(q + qi) + dq);  // q_out may alias q (own row only)
    st4(mu_out + mi, ldg4(mu + mi) + dm0);
    st4(mu_out + mi + NB_F, ldg4(mu + mi + NB_F) + dm1);
    st4(mu_out + mi + 2 * NB_F, ldg4(mu + mi + 2 * NB_F) + dm2);
}

// Backward, organised by SOURCE atom j.  For e in CSR row j (target j, source i = col[e]) the
// opposite edge e' = (j -> i) has the same filter row (W depends on d only) and unit vector -u_e,
// so every quantity of e' is available while streaming row j contiguously:
//   g_a_j += Wa * gq_i ;  g_b_j += Wb * (gmu_i . u') ;  g_c_j += Wc * sum_x gmu_i[x]*mu_j[x]
//   g_mu_j[x] += (Wc*c_j) * gmu_i[x]
//   dE/dd(e')   = sum_ch dWa*(a_j*gq_i) + dWb*(b_j*(gmu_i.u')) + dWc*(c_j*sum_x gmu_i[x] mu_j[x])
//   dE/du'(e')[x] = sum_ch (Wb*b_j) * gmu_i[x]
// The four edge scalars are warp-reduced and accumulated into egrad[e] (slot of e, values of e').
__global__ void __launch_bounds__(MSG_THREADS) k_painn_msg_bwd(const float* __restrict__ xh, const float* __restrict__ xh_bias,
                                                              const float* __restrict__ mu, const float* __restrict__ W,
                                                              const float* __restrict__ dW, const float* __restrict__ geom,
                                                              const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                              int n_atoms, const float* __restrict__ g_q, const float* __restrict__ g_mu,
                                                              float* __restrict__ g_xh, float* __restrict__ g_mu_in,
                                                              float* __restrict__ egrad) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * MSG_WARPS + (threadIdx.x >> 5);
    if (j >= n_atoms) return;
    const int c = lane * 4;
    const float* xj = xh + (size_t)j * (3 * NB_F) + c;
    const float4 a = ldg4(xj) + ldg4(xh_bias + c), b = ldg4(xj + NB_F) + ldg4(xh_bias + NB_F + c),
                 cc = ldg4(xj + 2 * NB_F) + ldg4(xh_bias + 2 * NB_F + c);
    const float* mj = mu + (size_t)j * (3 * NB_F) + c;
    const float4 m0 = ldg4(mj), m1 = ldg4(mj + NB_F), m2 = ldg4(mj + 2 * NB_F);
    float4 ga = f4(0.f), gb = f4(0.f), gc = f4(0.f), gm0 = f4(0.f), gm1 = f4(0.f), gm2 = f4(0.f);
    const int e0 = row_ptr[j], e1 = row_ptr[j + 1];
    for (int e = e0; e < e1; ++e) {
        const int i = __ldg(col + e);
        const float4 g = ldg4(geom + 4 * (size_t)e);  // u_e = (pos_i - pos_j)/d ; u' = -u_e
        const float* we = W + (size_t)e * (3 * NB_F) + c;
        const float4 wa = ldg4_stream(we), wb = ldg4_stream(we + NB_F), wc = ldg4_stream(we + 2 * NB_F);
        const float* dwe = dW + (size_t)e * (3 * NB_F) + c;
        const float4 da = ldg4_stream(dwe), db = ldg4_stream(dwe + NB_F), dc = ldg4_stream(dwe + 2 * NB_F);
        const float4 gq = ldg4(g_q + (size_t)i * NB_F + c);
        const float* gmi = g_mu + (size_t)i * (3 * NB_F) + c;
        const float4 h0 = ldg4(gmi), h1 = ldg4(gmi + NB_F), h2 = ldg4(gmi + 2 * NB_F);
        // t_b = gmu_i . u'   (per channel), t_c = sum_x gmu_i[x] * mu_j[x]
        float4 tb = h0 * (-g.x); fma4s(tb, h1, -g.y); fma4s(tb, h2, -g.z);
        float4 tc = h0 * m0; fma4(tc, h1, m1); fma4(tc, h2, m2);
        fma4(ga, wa, gq); fma4(gb, wb, tb); fma4(gc, wc, tc);
        const float4 pc = wc * cc;
        fma4(gm0, pc, h0); fma4(gm1, pc, h1); fma4(gm2, pc, h2);
        // edge scalars
        float4 sd = da * (a * gq); fma4(sd, db, b * tb); fma4(sd, dc, cc * tc);
        const float4 pb = wb * b;
        float gd = hsum4(sd), gu0 = hsum4(pb * h0), gu1 = hsum4(pb * h1), gu2 = hsum4(pb * h2);
        gd = warp_sum(gd); gu0 = warp_sum(gu0); gu1 = warp_sum(gu1); gu2 = warp_sum(gu2);
        if (lane == 0) {
            float4* slot = reinterpret_cast<float4*>(egrad + 4 * (size_t)e);
            float4 old = *slot;
            *slot = make_float4(old.x + gu0, old.y + gu1, old.z + gu2, old.w + gd);
        }
    }
    float* gx = g_xh + (size_t)j * (3 * NB_F) + c;
    st4(gx, ga); st4(gx + NB_F, gb); st4(gx + 2 * NB_F, gc);
    const float* gmj = g_mu + (size_t)j * (3 * NB_F) + c;
    float* go = g_mu_in + (size_t)j * (3 * NB_F) + c;
    st4(go, ldg4(gmj) + gm0); st4(go + NB_F, ldg4(gmj + NB_F) + gm1); st4(go + 2 * NB_F, ldg4(gmj + 2 * NB_F) + gm2);
}

// Forces from the accumulated edge gradients.  Slot e of row j holds, for the edge e' = (j -> i)
// with r' = pos_j - pos_i = -d u_e:  (dE/du'[3], dE/dd).  Chain rule through u' = r'/d, d = |r'|:
//   G(e) := dE/dr' = (gu - (gu.u') u')/d + gd u'
// pos_j receives +G(e) from its own row and -G(rev e) from the rows where it is the far end:
//   F_j = -dE/dpos_j = -sum_{e in row j} (G(e) - G(rev e))          (painn.py:135-146 autograd)
__device__ __forceinline__ float3 edge_G(const float4 eg, const float4 g) {
    const float ux = -g.x, uy = -g.y, uz = -g.z;  // u' of the opposite edge
    const float dot = eg.x * ux + eg.y * uy + eg.z * uz;
    const float inv = 1.0f / g.w;
    return make_float3((eg.x - dot * ux) * inv + eg.w * ux, (eg.y - dot * uy) * inv + eg.w * uy, (eg.z - dot * uz) * inv + eg.w * uz);
}

__global__ void __launch_bounds__(256) k_edge_forces(const float* __restrict__ egrad, const float* __restrict__ geom,
                                                    const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ rev, int n_atoms,
                                                    float* __restrict__ forces) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_atoms) return;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    for (int e = row_ptr[j]; e < row_ptr[j + 1]; ++e) {
        const int r = rev[e];
        const float3 g1 = edge_G(ldg4(egrad + 4 * (size_t)e), ldg4(geom + 4 * (size_t)e));
        const float3 g2 = edge_G(ldg4(egrad + 4 * (size_t)r), ldg4(geom + 4 * (size_t)r));
        fx -= g1.x - g2.x; fy -= g1.y - g2.y; fz -= g1.z - g2.z;
    }
    forces[3 * (size_t)j] = fx; forces[3 * (size_t)j + 1] = fy; forces[3 * (size_t)j + 2] = fz;
}

extern "C" int nb200_painn_msg_fwd(const float* xh, const float* xh_bias, const float* q, const float* mu, const float* W,
                                   const float* geom, const int32_t* row_ptr, const int32_t* col, int32_t n_atoms, float* q_out,
                                   float* mu_out, void* stream) {
    if (!xh || !xh_bias || !q || !mu || !W || !geom || !row_ptr || !col || !q_out || !mu_out || n_atoms < 0) return NB200_EINVAL;
    if (n_atoms == 0) return NB200_OK;
    k_painn_msg_fwd<<<(n_atoms + MSG_WARPS - 1) / MSG_WARPS, MSG_THREADS, 0, (cudaStream_t)stream>>>(xh, xh_bias, q, mu, W, geom, row_ptr,
                                                                                                     col, n_atoms, q_out, mu_out);
    return nb_check_launch();
}

extern "C" int nb200_painn_msg_bwd(const float* xh, const float* xh_bias, const float* mu, const float* W, const float* dW,
                                   const float* geom, const int32_t* row_ptr, const int32_t* col, int32_t n_atoms, const float* g_q,
                                   const float* g_mu, float* g_xh, float* g_mu_in, float* egrad, void* stream) {
    if (!xh || !xh_bias || !mu || !W || !dW || !geom || !row_ptr || !col || !g_q || !g_mu || !g_xh || !g_mu_in || !egrad || n_atoms < 0)
        return NB200_EINVAL;
    if (g_mu == g_mu_in) return NB200_EINVAL;
    if (n_atoms == 0) return NB200_OK;
    k_painn_msg_bwd<<<(n_atoms + MSG_WARPS - 1) / MSG_WARPS, MSG_THREADS, 0, (cudaStream_t)stream>>>(
        xh, xh_bias, mu, W, dW, geom, row_ptr, col, n_atoms, g_q, g_mu, g_xh, g_mu_in, egrad);
    return nb_check_launch();
}

extern "C" int nb200_edge_forces(const float* egrad, const float* geom, const int32_t* row_ptr, const int32_t* rev, int32_t n_atoms,
                                 float* forces, void* stream) {
    if (!egrad || !geom || !row_ptr || !rev || !forces || n_atoms < 0) return NB200_EINVAL;
    if (n_atoms == 0) return NB200_OK;
    k_edge_forces<<<(n_atoms + 255) / 256, 256, 0, (cudaStream_t)stream>>>(egrad, geom, row_ptr, rev, n_atoms, forces);
    return nb_check_launch();
}
