// graph.cu -- on-device neighbour build (CSR by target atom) for batches of small molecules.
//
// Replaces torch_cluster.radius_graph + the distance / unit-vector code of
// nablaDFT/painn_pyg/painn.py:411-423,306-321 (and ASENeighborList + PairwiseDistances for
// the schnetpack models).  One CTA per molecule: positions staged in shared memory, every
// thread owns target atoms and scans the molecule's sources in ascending order, so each CSR
// row is sorted by source and the whole layout is deterministic.
//
// HBM traffic: N*12 B read (x3 passes, L2-resident) + E*(4 col + 4 rev + 16 geom) B written.
#include "common.cuh"

#define NBR_THREADS 128
#define NBR_MAX_ATOMS 1024  // atoms per molecule staged in shared memory (12 KB)

// one expression for both passes so the count and the fill can never disagree on a pair
__device__ __forceinline__ float dist2(float dx, float dy, float dz) { return fmaf(dz, dz, fmaf(dy, dy, dx * dx)); }

__global__ void __launch_bounds__(NBR_THREADS) k_nbr_count(const float* __restrict__ pos, const int32_t* __restrict__ mol_ptr,
                                                          float cutoff2, int32_t max_neighbors, int32_t* __restrict__ deg,
                                                          int32_t* __restrict__ status) {
    __shared__ float sp[NBR_MAX_ATOMS * 3];
    const int m = blockIdx.x;
    const int a0 = mol_ptr[m], n = mol_ptr[m + 1] - a0;
    if (n > NBR_MAX_ATOMS) {
        if (threadIdx.x == 0) atomicMin(&status[1], NB200_EUNSUPPORTED);
        return;
    }
    for (int t = threadIdx.x; t < n * 3; t += NBR_THREADS) sp[t] = pos[(size_t)a0 * 3 + t];
    __syncthreads();
    int local_max = 0, local_iso = 0;
    for (int i = threadIdx.x; i < n; i += NBR_THREADS) {
        const float xi = sp[3 * i], yi = sp[3 * i + 1], zi = sp[3 * i + 2];
        int c = 0;
        for (int j = 0; j < n; ++j) {
            const float dx = sp[3 * j] - xi, dy = sp[3 * j + 1] - yi, dz = sp[3 * j + 2] - zi;
            const float d2 = dist2(dx, dy, dz);
            c += (d2 < cutoff2 && j != i) ? 1 : 0;
        }
        deg[a0 + i] = c;
        local_max = max(local_max, c);
        local_iso += (c == 0);
    }
    if (local_max > 0) atomicMax(&status[2], local_max);
    if (local_iso > 0) atomicAdd(&status[3], local_iso);
    if (local_max > max_neighbors) atomicMin(&status[1], NB200_ENEIGHBORS);
}

// single-CTA exclusive scan: deg[N] -> row_ptr[N+1]; publishes the edge count.
__global__ void __launch_bounds__(1024) k_nbr_scan(const int32_t* __restrict__ deg, int32_t n_atoms, int32_t e_cap,
                                                  int32_t* __restrict__ row_ptr, int32_t* __restrict__ status) {
    __shared__ int32_t warp_tot[32];
    const int tid = threadIdx.x;
    const int chunk = (n_atoms + 1023) / 1024;
    const int lo = min(tid * chunk, n_atoms), hi = min(lo + chunk, n_atoms);
    int32_t s = 0;
    for (int i = lo; i < hi; ++i) s += deg[i];
    // inclusive scan of per-thread sums
    int32_t v = s;
    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    if (w == 0) {
        int32_t t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    int32_t base = v - s + (w > 0 ? warp_tot[w - 1] : 0);
    for (int i = lo; i < hi; ++i) {
        row_ptr[i] = base;
        base += deg[i];
    }
    if (tid == 1023) {
        const int32_t total = warp_tot[31];
        row_ptr[n_atoms] = total;
        status[0] = total;
        if (total > e_cap) atomicMin(&status[1], NB200_ECAPACITY);
    }
}

__global__ void __launch_bounds__(NBR_THREADS) k_nbr_fill(const float* __restrict__ pos, const int32_t* __restrict__ mol_ptr,
                                                         float cutoff2, int32_t* __restrict__ row_ptr,
                                                         int32_t* __restrict__ col, int32_t* __restrict__ rev,
                                                         float* __restrict__ geom, const int32_t* __restrict__ status) {
    __shared__ float sp[NBR_MAX_ATOMS * 3];
    const int m = blockIdx.x;
    const int a0 = mol_ptr[m], n = mol_ptr[m + 1] - a0;
    if (status[1] != 0) {
        // capacity / neighbour-cap error: write no edges and make every CSR row empty, so the
        // downstream kernels touch nothing out of bounds; the host reads the flag with the results
        for (int t = threadIdx.x; t <= n; t += NBR_THREADS) row_ptr[a0 + t] = 0;
        return;
    }
    for (int t = threadIdx.x; t < n * 3; t += NBR_THREADS) sp[t] = pos[(size_t)a0 * 3 + t];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NBR_THREADS) {
        const float xi = sp[3 * i], yi = sp[3 * i + 1], zi = sp[3 * i + 2];
        int e = row_ptr[a0 + i];
        for (int j = 0; j < n; ++j) {
            // r = pos[source j] - pos[target i]   (painn.py:419 distance_vec = pos[j] - pos[i])
            const float dx = sp[3 * j] - xi, dy = sp[3 * j + 1] - yi, dz = sp[3 * j + 2] - zi;
            const float d2 = dist2(dx, dy, dz);
            if (d2 < cutoff2 && j != i) {
                const float d = sqrtf(d2);
                // painn.py:319-321: divide by (d + 1e-6*[d ~ 0]); atoms never coincide, keep the guard
                const float inv = 1.0f / (d + (d < 1e-6f ? 1e-6f : 0.0f));
                col[e] = a0 + j;
                st4(geom + 4 * (size_t)e, make_float4(dx * inv, dy * inv, dz * inv, d));
                ++e;
            }
        }
    }
    __syncthreads();  // the CTA's own global writes of col[] are visible to it after the barrier
    for (int i = threadIdx.x; i < n; i += NBR_THREADS) {
        const int gi = a0 + i;
        for (int e = row_ptr[gi]; e < row_ptr[gi + 1]; ++e) {
            const int gj = col[e];
            int lo = row_ptr[gj], hi = row_ptr[gj + 1] - 1;  // find gi in row gj (sorted ascending)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (col[mid] < gi) lo = mid + 1; else hi = mid;
            }
            rev[e] = lo;
        }
    }
}

extern "C" int nb200_neighbor_build(const float* pos, const int32_t* mol_ptr, int32_t n_mol, int32_t n_atoms, float cutoff,
                                    int32_t max_neighbors, int32_t e_cap, int32_t* row_ptr, int32_t* col, int32_t* rev,
                                    float* geom, int32_t* deg_scratch, int32_t* status, void* stream) {
    if (!pos || !mol_ptr || !row_ptr || !col || !rev || !geom || !deg_scratch || !status) return NB200_EINVAL;
    if (n_mol < 0 || n_atoms < 0 || e_cap < 0 || !(cutoff > 0.f)) return NB200_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), s) != cudaSuccess) return nb_check_launch();
    if (n_mol == 0 || n_atoms == 0) return cudaMemsetAsync(row_ptr, 0, (size_t)(n_atoms + 1) * 4, s) == cudaSuccess ? NB200_OK : nb_check_launch();
    const float c2 = cutoff * cutoff;
    k_nbr_count<<<n_mol, NBR_THREADS, 0, s>>>(pos, mol_ptr, c2, max_neighbors, deg_scratch, status);
    k_nbr_scan<<<1, 1024, 0, s>>>(deg_scratch, n_atoms, e_cap, row_ptr, status);
    k_nbr_fill<<<n_mol, NBR_THREADS, 0, s>>>(pos, mol_ptr, c2, row_ptr, col, rev, geom, status);
    return nb_check_launch();
}
