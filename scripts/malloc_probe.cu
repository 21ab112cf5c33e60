// malloc_probe.cu — what do cudaMalloc / cudaFree of the PDLP handle's buffers cost?  (C4: 6 matrix arrays of 80-160 MB,
// ~20 vectors of 16-32 MB.)  Build + run on the GPU box:
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a scripts/malloc_probe.cu -o scripts/_build/malloc_probe && scripts/_build/malloc_probe
#include <chrono>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaFree(0);
    const size_t sizes[] = {16u << 20, 32u << 20, 80u << 20, 160u << 20, 240u << 20};
    for (int round = 0; round < 2; ++round)
        for (size_t s : sizes) {
            std::vector<void*> p(8);
            double t0 = now();
            for (auto& q : p) cudaMalloc(&q, s);
            double t1 = now();
            for (auto& q : p) cudaMemset(q, 1, s);
            cudaDeviceSynchronize();
            double t2 = now();
            for (auto& q : p) cudaFree(q);
            double t3 = now();
            printf("{\"MB\": %zu, \"round\": %d, \"malloc_ms_each\": %.3f, \"free_ms_each\": %.3f}\n", s >> 20, round, (t1 - t0) / 8, (t3 - t2) / 8);
        }
    // one slab instead: 40 buffers' worth (1.2 GB)
    void* slab;
    double t0 = now();
    cudaMalloc(&slab, (size_t)1200 << 20);
    double t1 = now();
    cudaMemset(slab, 1, (size_t)1200 << 20);
    cudaDeviceSynchronize();
    double t2 = now();
    cudaFree(slab);
    double t3 = now();
    printf("{\"MB\": 1200, \"slab\": 1, \"malloc_ms\": %.3f, \"free_ms\": %.3f}\n", t1 - t0, t3 - t2);
    return 0;
}
