// primitives.cuh — device-wide exclusive scan and a STABLE LSD radix sort (64-bit keys, 32-bit
// payload), hand-written for sm_100a.  Both are HBM-bound streaming passes; they back
//   * CSR assembly (assemble.cu): stable sort of (row,col) keys keeps emission order inside a key,
//     which is what makes the ordered duplicate reduction bit-exact;
//   * the device transpose CSR -> CSC used by the PDLP A'y kernel (pdlp.cu).
#pragma once
#include "common.cuh"

namespace elp {

// ------------------------------------------------------------------------------------------------
// exclusive scan over uint32 (in place).  Three-phase, recursive over block sums.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t s[SCAN_TILE];
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        uint32_t i = base + it * SCAN_THREADS + threadIdx.x;
        s[it * SCAN_THREADS + threadIdx.x] = i < n ? data[i] : 0u;
    }
    __syncthreads();
    // each thread owns SCAN_ITEMS consecutive elements
    uint32_t loc[SCAN_ITEMS];
    uint32_t tot = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        loc[k] = tot;
        tot += s[threadIdx.x * SCAN_ITEMS + k];
    }
    // exclusive scan of thread totals across the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += warp_tot[w];
    const uint32_t thread_off = warp_off + inc - tot;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) s[threadIdx.x * SCAN_ITEMS + k] = thread_off + loc[k];
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        uint32_t i = base + it * SCAN_THREADS + threadIdx.x;
        if (i < n) data[i] = s[it * SCAN_THREADS + threadIdx.x];
    }
    if (threadIdx.x == SCAN_THREADS - 1 && tile_sums) tile_sums[blockIdx.x] = thread_off + tot;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ tile_offs) {
    const uint32_t off = tile_offs[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_TILE;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        uint32_t i = base + it * SCAN_THREADS + threadIdx.x;
        if (i < n) data[i] += off;
    }
}

// Workspace big enough for scans of up to `max_n` elements.
struct ScanWorkspace {
    DevBuf<uint32_t> lvl1, lvl2, lvl3;
    void reserve(size_t max_n) {
        size_t n1 = (max_n + SCAN_TILE - 1) / SCAN_TILE;
        size_t n2 = (n1 + SCAN_TILE - 1) / SCAN_TILE;
        size_t n3 = (n2 + SCAN_TILE - 1) / SCAN_TILE;
        if (lvl1.n < n1 + 1) lvl1.alloc(n1 + 1);
        if (lvl2.n < n2 + 1) lvl2.alloc(n2 + 1);
        if (lvl3.n < n3 + 1) lvl3.alloc(n3 + 1);
    }
};

inline void exclusive_scan_u32(uint32_t* d, size_t n, ScanWorkspace& ws, cudaStream_t st) {
    if (n == 0) return;
    ws.reserve(n);
    uint32_t* lv[3] = {ws.lvl1.p, ws.lvl2.p, ws.lvl3.p};
    size_t cnt[4];
    cnt[0] = n;
    for (int l = 0; l < 3; ++l) cnt[l + 1] = (cnt[l] + SCAN_TILE - 1) / SCAN_TILE;
    ELP_REQUIRE(cnt[3] == 1, "exclusive_scan_u32: input too large (%zu)", n);
    uint32_t* cur[3] = {d, lv[0], lv[1]};
    // down-sweep: tile scans, collecting tile sums into the next level
    int depth = 0;
    for (int l = 0; l < 3; ++l) {
        ELP_LAUNCH(scan_tile_kernel, (unsigned)cnt[l + 1], SCAN_THREADS, 0, st, cur[l], (uint32_t)cnt[l], lv[l]);
        depth = l;
        if (cnt[l + 1] == 1) break;
    }
    // up-sweep: add scanned tile offsets back
    for (int l = depth - 1; l >= 0; --l)
        ELP_LAUNCH(scan_add_kernel, (unsigned)cnt[l + 1], SCAN_THREADS, 0, st, cur[l], (uint32_t)cnt[l], lv[l]);
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass.
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_RADIX = 256;

static __global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ counts,
               uint32_t ntiles) {
    __shared__ uint32_t h[RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        uint32_t i = base + it * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];   // digit-major
}

static __global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                  uint64_t* __restrict__ kout, uint32_t* __restrict__ vout, uint32_t n, int shift,
                  const uint32_t* __restrict__ offsets, uint32_t ntiles) {
    __shared__ uint32_t whist[RS_WARPS][RS_RADIX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    // warp w owns the contiguous slice [w*RS_ITEMS*32, (w+1)*RS_ITEMS*32) of the tile; within the
    // slice the order is (it, lane) == ascending global index, so ranks below are stable.
    const uint32_t wbase = blockIdx.x * RS_TILE + warp * (RS_ITEMS * 32);
    uint64_t k[RS_ITEMS];
    uint32_t v[RS_ITEMS], r[RS_ITEMS];
    int dg[RS_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const uint32_t i = wbase + it * 32 + lane;
        const bool valid = i < n;
        k[it] = valid ? kin[i] : 0ull;
        v[it] = valid ? vin[i] : 0u;
        const int d = valid ? (int)((uint32_t)(k[it] >> shift) & 255u) : RS_RADIX;   // sentinel for padding
        dg[it] = d;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t pre = 0;
        if (valid) pre = whist[warp][d];
        __syncwarp();
        r[it] = pre + __popc(peers & lt_mask);
        if (valid && lane == (__ffs(peers) - 1)) whist[warp][d] = pre + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: exclusive prefix over warps + the tile's global offset for digit d
        const int d = threadIdx.x;
        uint32_t run = offsets[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = whist[w][d];
            whist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        if (dg[it] < RS_RADIX) {
            const uint32_t pos = whist[warp][dg[it]] + r[it];
            kout[pos] = k[it];
            vout[pos] = v[it];
        }
    }
}

struct RadixSortWorkspace {
    DevBuf<uint64_t> kalt;
    DevBuf<uint32_t> valt;
    DevBuf<uint32_t> counts;
    ScanWorkspace scan;
};

// Sorts (keys, vals) of length n by the low `nbits` bits of the key, stably.  On return the sorted
// data is in (keys, vals) again.
inline void radix_sort_pairs(uint64_t* keys, uint32_t* vals, size_t n, int nbits, RadixSortWorkspace& ws,
                             cudaStream_t st) {
    if (n <= 1 || nbits <= 0) return;
    ELP_REQUIRE(n < 0xffffffffull, "radix_sort_pairs: too many items (%zu)", n);
    const uint32_t ntiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    if (ws.kalt.n < n) ws.kalt.alloc(n);
    if (ws.valt.n < n) ws.valt.alloc(n);
    if (ws.counts.n < (size_t)ntiles * RS_RADIX) ws.counts.alloc((size_t)ntiles * RS_RADIX);
    uint64_t* kin = keys;  uint32_t* vin = vals;
    uint64_t* kout = ws.kalt.p;  uint32_t* vout = ws.valt.p;
    const int passes = (nbits + 7) / 8;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        ELP_LAUNCH(rs_hist_kernel, ntiles, RS_THREADS, 0, st, kin, (uint32_t)n, shift, ws.counts.p, ntiles);
        exclusive_scan_u32(ws.counts.p, (size_t)ntiles * RS_RADIX, ws.scan, st);
        ELP_LAUNCH(rs_scatter_kernel, ntiles, RS_THREADS, 0, st, kin, vin, kout, vout, (uint32_t)n, shift,
                   ws.counts.p, ntiles);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    if (kin != keys) {
        ELP_CUDA(cudaMemcpyAsync(keys, kin, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        ELP_CUDA(cudaMemcpyAsync(vals, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
}

inline int bit_length_u64(uint64_t v) {
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

}  // namespace elp
