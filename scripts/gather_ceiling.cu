// gather_ceiling.cu — how fast can a B200 do what the SpMV kernels cannot avoid?
// 20 M random 8-byte gathers from an L2-resident vector (32 MB / 16 MB), index stream read coalesced, one 8-byte
// result per 10 / 5 gathers written back: the access pattern of A.x-bar / A'.y on the C4 matrix with everything else
// (matrix values, TMA staging, epilogue) removed.  Build + run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/gather_ceiling.cu -o /tmp/gather_ceiling && /tmp/gather_ceiling
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

template <int PER>
__global__ void __launch_bounds__(256) gather_kernel(const int* __restrict__ idx, const double* __restrict__ vec,
                                                     double* __restrict__ out, int nrows) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        int c[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) c[u] = __ldg(idx + (size_t)u * nrows + r);      // coalesced index stream
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < PER; ++u) s += __ldg(vec + c[u]);
        out[r] = s;
    }
}

// the dual question: 20 M random fire-and-forget fp64 reductions (RED.ADD) into the same vector
template <int PER>
__global__ void __launch_bounds__(256) scatter_kernel(const int* __restrict__ idx, double* __restrict__ vec, int nrows) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        int c[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) c[u] = __ldg(idx + (size_t)u * nrows + r);
#pragma unroll
        for (int u = 0; u < PER; ++u) atomicAdd(vec + c[u], 1.0);
    }
}

template <int PER>
static void run(int nvec, int nrows) {
    std::vector<int> h((size_t)PER * nrows);
    uint64_t st = 88172645463325252ull;
    for (auto& v : h) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = (int)(st % (uint64_t)nvec); }
    int* idx; double *vec, *out;
    cudaMalloc(&idx, h.size() * 4); cudaMalloc(&vec, (size_t)nvec * 8); cudaMalloc(&out, (size_t)nrows * 8);
    cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(vec, 0, (size_t)nvec * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        for (int i = 0; i < 3; ++i) gather_kernel<PER><<<grid, 256>>>(idx, vec, out, nrows);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) gather_kernel<PER><<<grid, 256>>>(idx, vec, out, nrows);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
        const double g = (double)PER * nrows;
        printf("{\"gathers\": %.0f, \"per_row\": %d, \"vec_MB\": %.0f, \"grid\": %d, \"ms\": %.4f, \"Ggather_per_s\": %.1f}\n", g, PER,
               nvec * 8 / 1e6, grid, ms, g / ms / 1e6);
    }
    for (int grid : {148 * 8, 148 * 32}) {
        for (int i = 0; i < 3; ++i) scatter_kernel<PER><<<grid, 256>>>(idx, vec, nrows);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) scatter_kernel<PER><<<grid, 256>>>(idx, vec, nrows);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
        const double g = (double)PER * nrows;
        printf("{\"red_adds\": %.0f, \"per_row\": %d, \"vec_MB\": %.0f, \"grid\": %d, \"ms\": %.4f, \"Gred_per_s\": %.1f}\n", g, PER,
               nvec * 8 / 1e6, grid, ms, g / ms / 1e6);
    }
    cudaFree(idx); cudaFree(vec); cudaFree(out);
}

int main() {
    run<10>(4000000, 2000000);   // A.x-bar on C4: 2 M rows x 10 gathers from a 32 MB vector
    run<5>(2000000, 4000000);    // A'.y   on C4: 4 M rows x 5 gathers from a 16 MB vector
    return 0;
}
