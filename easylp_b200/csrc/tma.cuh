// tma.cuh — 1-D TMA bulk copies (cp.async.bulk -> mbarrier complete_tx) and mbarrier helpers, raw PTX.
// Used by the batched simplex (tableau staging) and the PDLP SpMV (matrix stream staging).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace elp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spins until the phase with the given parity completes.  A bounded spin: a copy that never lands
// (bad pointer / size) traps the kernel instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    } while (!done);
}
// global -> shared bulk copy; bytes and both addresses must be multiples of 16
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same, with an L2 eviction-priority hint (policy from l2_policy_evict_first / l2_policy_evict_last)
__device__ __forceinline__ void tma_load_1d_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ldg_hint(const double* ptr, uint64_t policy) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(ptr), "l"(policy));
    return v;
}
__device__ __forceinline__ void stg_hint(double* ptr, double v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(ptr), "d"(v), "l"(policy) : "memory");
}
// L2 residency hints of one SpMV launch (bit flags; spmv.cuh): what the streams that are read once and the vectors
// that the NEXT kernel gathers from ask of the L2.
struct L2Hints {
    uint64_t first, last;
    int flags;      // 1: operand streams evict-first, 2: non-gathered outputs evict-first, 4: gathered outputs evict-last,
                    // 8: gathers evict-last
};
__device__ __forceinline__ void st_out(double* ptr, double v, const L2Hints& h, bool gathered_next) {
    if (gathered_next) { if (h.flags & 4) stg_hint(ptr, v, h.last); else *ptr = v; }
    else { if (h.flags & 2) stg_hint(ptr, v, h.first); else *ptr = v; }
}
__device__ __forceinline__ bool tma_ok(const void* src, size_t bytes) {
    return bytes > 0 && (bytes & 15) == 0 && (((uintptr_t)src) & 15) == 0;
}

}  // namespace elp
