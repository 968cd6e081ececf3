// lsu_rate.cu -- per-SM global store / load throughput with (almost) no L1: the question behind the fused node kernels' epilogues.
// 512 threads per CTA, 200 KB of dynamic shared memory (as the node kernels: the L1 carve-out is what is left), one CTA per SM.
//   mode 0: every warp stores 128-byte rows (lane = 4 bytes, row stride 3 KB like VW)     mode 1: same with 16-byte lanes (512 B per warp)
//   mode 2: loads, 128 B per warp-instruction, 16 independent loads in flight per thread   mode 3: loads, 512 B per warp-instruction
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(512, 1) k_lsu(int mode, int reps, float* buf, size_t per_cta_floats, unsigned long long* cyc, float* sink) {
    extern __shared__ float sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* base = buf + (size_t)blockIdx.x * per_cta_floats;
    if (tid == 0) sm[0] = 0.f;
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    // one "tile" = 128 rows x 128 floats (64 KB), row stride 768 floats (3 KB, like VW); 4 column blocks cycled; no integer division in the loop
    for (int r = 0; r < reps; ++r) {
        float* col = base + (r & 3) * 128;
        if (mode == 0) {        // 16 warps x 8 rows, lane = 4 bytes: 128 B per warp-instruction
#pragma unroll
            for (int j = 0; j < 8; ++j) col[(size_t)(warp * 8 + j) * 768 + lane] = (float)j;
        } else if (mode == 1) { // 16 warps x 2 row-quads... 512 B per warp-instruction: lane = 16 bytes, a warp covers one 512-byte row
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(base + (size_t)(warp * 8 + j) * 768 + (r & 1) * 128 + lane * 4) = make_float4(j, j, j, j);
        } else if (mode == 2) {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldcg(col + (size_t)(warp * 8 + j) * 768 + lane);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += t[j];
        } else {
            float4 t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldcg(reinterpret_cast<const float4*>(base + (size_t)(warp * 8 + j) * 768 + (r & 1) * 128 + lane * 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += t[j].x + t[j].y + t[j].z + t[j].w;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (acc == 1234.5f) sink[0] = acc;
    if (tid == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

int main() {
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(k_lsu, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const size_t per_cta = 128 * 768;  // floats: 384 KB region per CTA (rows of 3 KB)
    float *buf, *sink;
    CK(cudaMalloc(&buf, 148 * per_cta * 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, 148 * per_cta * 4));
    unsigned long long* cyc;
    CK(cudaMalloc(&cyc, 148 * 8));
    const int reps = 200;
    const char* names[] = {"store 128 B / warp-instr", "store 512 B / warp-instr", "load  128 B / warp-instr", "load  512 B / warp-instr"};
    for (int grid : {1, 76, 148}) for (int mode = 0; mode < 4; ++mode) {
        for (int it = 0; it < 2; ++it) { k_lsu<<<grid, 512, smem>>>(mode, reps, buf, per_cta, cyc, sink); CK(cudaDeviceSynchronize()); }
        std::vector<unsigned long long> c(grid);
        CK(cudaMemcpy(c.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
        double s = 0; for (auto v : c) s += (double)v; s /= grid;
        { const double bytes = (mode == 0 || mode == 2) ? 16.0 * 8 * 128 : 16.0 * 8 * 512; printf("grid %3d %s: %7.1f cycles per pass = %5.1f B/clk per SM\n", grid, names[mode], s / reps, bytes * reps / s); }
    }
    return 0;
}
