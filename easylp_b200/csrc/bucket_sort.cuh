// bucket_sort.cuh — assembly of UNORDERED term streams without a multi-pass global sort.
//
// The stable LSD radix sort of primitives.cuh moves every (key, index) pair through HBM once per 8-bit digit: six passes
// for the 43-bit keys of a 2M x 4M matrix, each with scattered 8-byte writes (ncu: 360 us per pass for 21 M terms, 3.4 ms
// for the whole assembly = 2.6 % of the HBM roofline).  The CSR target has structure the sort ignores: rows are short.
//   A. ONE split by row range: a bucket is `rpb` consecutive rows, sized so that its terms fit one CTA's shared memory.
//      Bucket sizes: every CTA of a persistent grid histograms its chunk of the stream in shared memory and adds its
//      non-empty bins to the global counts (a RED per term took 240-290 us for 21 M terms, this takes ~70).  After a scan
//      of the sizes every term takes the next free slot of its bucket (one returning atomic per term) and stores a
//      record {row-in-bucket | col, stream index, value}, padded to one full 32-byte sector and written by one 256-bit
//      store.  The order inside a bucket is whatever the atomics gave.
//      Measured alternatives for the placement (21 M terms, 19.6 K buckets): per-(bucket, CTA) ranges from a scanned count
//      matrix with shared-memory cursors, no global atomics — 0.88 ms with 16-byte records (every store a partial-sector
//      write: the L2 fetched the other half first, ncu counted 2.5 GB of DRAM reads), 0.67 ms with 32-byte records
//      (5.8 M write fronts, each touched too rarely for its lines to complete in L2: 21 M random 32-byte DRAM writes);
//      sweeping the buckets in L2-sized ranges cost ~100 us per sweep and gained nothing.  With ONE cursor per bucket
//      the records of a bucket arrive back to back from all SMs, lines complete within microseconds: 0.39 ms.
//   B. One CTA per bucket, everything in shared memory: counting sort by row, then every term ranks itself inside its
//      row by (col, stream index) — rows are short, so the quadratic count is cheap — which restores EMISSION ORDER inside
//      equal (row, col) whatever phase A did; heads of equal-key runs fold left to right with plain fp64 adds (the
//      Reduce(`+`) order of /root/reference/R/methods.R:248-250), exact zeros are dropped, survivors are written
//      compacted to the bucket's own range together with per-row counts.
//   C. Scan the buckets' survivor counts; one CTA per bucket copies its survivors to their final place and writes its
//      rows' row_ptr entries.
// Traffic ~ 4 + 48 + 32 + 12 + 24 bytes per term instead of ~8 x 24; the result is bit-identical to the sorted path
// (tests/test_gpu_kernels.py runs both on the same streams).  A bucket that does not fit (one very long row, a skewed
// stream) makes the caller fall back to the radix sort.
#pragma once
#include "common.cuh"
#include "primitives.cuh"
#include <cmath>

namespace elp {

constexpr int BK_THREADS = 256;              // also the most rows a bucket may span: one row counter per thread
constexpr int BK_CAP = 1536;                 // terms per bucket: 32 B x 1536 = 48 KB of shared memory, four CTAs per SM
constexpr int BK_SPLIT_THREADS = 512;        // histogram of phase A: two CTAs per SM, BK_SPLIT_U loads in flight per thread
constexpr int BK_SPLIT_U = 8;
constexpr int BK_PER = BK_CAP / BK_THREADS;  // consecutive terms a thread owns in the compaction
constexpr uint32_t BK_MAX_BUCKETS = 1u << 22;

// A record is padded to ONE 32-byte sector and stored with one 256-bit instruction (sm_100: STG.E.ENL2.256).  With 16-byte
// records every store of the split was a partial-sector write to a random address and the L2 fetched the other half
// first: ncu counted 2.5 GB of DRAM reads for 0.8 GB of input, the split took 0.9 ms of the 1.5 ms.
struct __align__(32) BkRec {
    uint32_t key;      // (row - first row of the bucket) << colbits | col
    uint32_t idx;      // position in the caller's stream
    double val;
    unsigned long long pad[2];
};
__device__ __forceinline__ void bk_store_rec(BkRec* dst, uint32_t key, uint32_t idx, double val) {
    const unsigned long long a = ((unsigned long long)idx << 32) | key, b = (unsigned long long)__double_as_longlong(val);
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %3};" ::"l"(dst), "l"(a), "l"(b), "l"(0ull) : "memory");
}

// exclusive scan of one value per thread; every thread gets the block total too.  `ws` needs blockDim/32 + 1 words.
__device__ __forceinline__ uint32_t bk_block_scan(uint32_t v, uint32_t* ws, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < nw ? ws[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < nw) ws[lane] = wi - w;
        if (lane == 31) ws[nw] = wi;
    }
    __syncthreads();
    const uint32_t r = ws[warp] + inc - v;
    total = ws[nw];
    __syncthreads();                    // ws may be reused right away
    return r;
}

// A1: bucket sizes
static __global__ void bk_count(const int32_t* __restrict__ row, uint32_t T, uint32_t rpb, uint32_t m, uint32_t* __restrict__ counts,
                                int* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint32_t r = (uint32_t)row[i];
    if (r < m) atomicAdd(counts + r / rpb, 1u);
    else *bad = 1;                                  // reported by the host; never counted, never placed
}

// A2 / C1: single CTA, exclusive scan of nb values -> start[0..nb] (and a copy the scatter uses as its cursors);
// stats[0] = largest value, stats[1] = sum
static __global__ void __launch_bounds__(1024)
bk_offsets(const uint32_t* __restrict__ counts, uint32_t nb, uint32_t* __restrict__ start, uint32_t* __restrict__ cursor,
           uint32_t* __restrict__ stats) {
    __shared__ uint32_t ws[33];
    __shared__ uint32_t smax;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    uint32_t carry = 0, mx = 0;
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? counts[i] : 0u;
        mx = max(mx, v);
        uint32_t tot;
        const uint32_t ex = carry + bk_block_scan(v, ws, tot);
        if (i < nb) {
            start[i] = ex;
            if (cursor) cursor[i] = ex;
        }
        carry += tot;
    }
    atomicMax(&smax, mx);
    __syncthreads();
    if (threadIdx.x == 0) { start[nb] = carry; stats[0] = smax; stats[1] = carry; }
}

// A3: records into their bucket's range
static __global__ void bk_scatter(const int32_t* __restrict__ row, const int32_t* __restrict__ col, const double* __restrict__ val,
                                  uint32_t T, uint32_t rpb, int colbits, uint32_t m, uint32_t n, uint32_t* __restrict__ cursor,
                                  BkRec* __restrict__ rec, int* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint32_t r = (uint32_t)row[i], b = r / rpb;
    if (r >= m) return;
    const uint32_t pos = atomicAdd(cursor + b, 1u);
    uint32_t c = (uint32_t)col[i];
    if (c >= n) { *bad = 1; c = 0; }
    bk_store_rec(rec + pos, ((r - b * rpb) << colbits) | c, i, val[i]);
}

// A1 (buckets fit a shared-memory histogram): CTA g counts terms [g * chunk, (g + 1) * chunk) in shared memory and adds
// its non-empty bins to the global counts
static __global__ void __launch_bounds__(BK_SPLIT_THREADS)
bk_hist_cta(const int32_t* __restrict__ row, uint32_t T, uint32_t chunk, uint32_t rpb, uint32_t nb, uint32_t m,
            uint32_t* __restrict__ counts, int* __restrict__ bad) {
    extern __shared__ uint32_t bk_h[];
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) bk_h[b] = 0;
    __syncthreads();
    const uint32_t lo = blockIdx.x * chunk, hi = (uint32_t)min((uint64_t)T, (uint64_t)lo + chunk);
    for (uint64_t base = lo; base < hi; base += (uint64_t)BK_SPLIT_THREADS * BK_SPLIT_U) {
        uint32_t r[BK_SPLIT_U];
#pragma unroll
        for (int k = 0; k < BK_SPLIT_U; ++k) {
            const uint64_t i = base + (uint64_t)k * BK_SPLIT_THREADS + threadIdx.x;
            r[k] = i < hi ? (uint32_t)row[i] : 0xffffffffu;
        }
#pragma unroll
        for (int k = 0; k < BK_SPLIT_U; ++k) {
            const uint64_t i = base + (uint64_t)k * BK_SPLIT_THREADS + threadIdx.x;
            if (i >= hi) continue;
            if (r[k] < m) atomicAdd(&bk_h[r[k] / rpb], 1u);
            else *bad = 1;                          // reported by the host; never counted, never placed
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
        const uint32_t c = bk_h[b];
        if (c) atomicAdd(counts + b, c);
    }
}

// B: sort and fold one bucket, rec[start[b] .. start[b + 1]), in shared memory
static __global__ void __launch_bounds__(BK_THREADS, 4)
bk_sort_fold(const BkRec* __restrict__ rec, const uint32_t* __restrict__ start, uint32_t rpb, int colbits, uint32_t m, int32_t* __restrict__ tcol, double* __restrict__ tval,
             uint32_t* __restrict__ bnnz, uint32_t* __restrict__ rowcnt, uint32_t* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char bk_smem[];
    uint32_t* K1 = reinterpret_cast<uint32_t*>(bk_smem);          // keys as loaded; later the sorted keys
    uint32_t* I1 = K1 + BK_CAP;                                   // stream indices as loaded; later the keep flags
    unsigned long long* KI2 = reinterpret_cast<unsigned long long*>(I1 + BK_CAP);     // key << 32 | index, grouped by row
    double* V1 = reinterpret_cast<double*>(KI2 + BK_CAP);         // values as loaded; later the sorted values
    double* V2 = V1 + BK_CAP;                                     // values grouped by row; later the folded sums
    __shared__ uint32_t rcnt[BK_THREADS], roff[BK_THREADS + 1], ws[BK_THREADS / 32 + 1];
    const uint32_t b = blockIdx.x, tid = threadIdx.x;
    const uint32_t s0 = start[b], cnt = start[b + 1] - s0;
    const uint32_t row0 = b * rpb;
    const uint32_t nr = min(rpb, m - row0);
    const uint32_t colmask = colbits >= 32 ? 0xffffffffu : ((1u << colbits) - 1u);
    if (cnt > (uint32_t)BK_CAP) {             // does not fit: the caller falls back to the radix sort
        if (tid == 0) { atomicMax(overflow, cnt); bnnz[b] = 0; }
        if (tid < nr) rowcnt[row0 + tid] = 0;
        return;
    }
    rcnt[tid] = 0;
    __syncthreads();
    // 1. load, count per row
    {
        int4 x[BK_PER];                          // all loads in flight before the first is used (a bucket comes from HBM)
        const BkRec* src = rec + s0;
#pragma unroll
        for (int k = 0; k < BK_PER; ++k) {
            const uint32_t e = tid + (uint32_t)k * BK_THREADS;
            if (e < cnt) x[k] = __ldcs(reinterpret_cast<const int4*>(src + e));      // the first half of the sector
        }
#pragma unroll
        for (int k = 0; k < BK_PER; ++k) {
            const uint32_t e = tid + (uint32_t)k * BK_THREADS;
            if (e < cnt) {
                const uint32_t key = (uint32_t)x[k].x;
                K1[e] = key; I1[e] = (uint32_t)x[k].y;
                V1[e] = __hiloint2double(x[k].w, x[k].z);
                atomicAdd(&rcnt[key >> colbits], 1u);
            }
        }
    }
    __syncthreads();
    // 2. row offsets
    {
        uint32_t tot;
        const uint32_t ex = bk_block_scan(rcnt[tid], ws, tot);
        roff[tid] = ex;
        if (tid == 0) roff[BK_THREADS] = tot;
        rcnt[tid] = ex;
    }
    __syncthreads();
    // 3. counting sort by row (order inside a row: arbitrary)
    for (uint32_t e = tid; e < cnt; e += BK_THREADS) {
        const uint32_t k = K1[e];
        const uint32_t p = atomicAdd(&rcnt[k >> colbits], 1u);
        KI2[p] = ((unsigned long long)k << 32) | I1[e];
        V2[p] = V1[e];
    }
    __syncthreads();
    // 4. every term ranks itself inside its row by (col, stream index): ascending columns, emission order inside a column
    for (uint32_t p = tid; p < cnt; p += BK_THREADS) {
        const unsigned long long ki = KI2[p];
        const uint32_t k = (uint32_t)(ki >> 32);
        const uint32_t r = k >> colbits;
        const uint32_t lo = roff[r], hi = roff[r + 1];
        uint32_t rank = 0;
        for (uint32_t q = lo; q < hi; ++q) rank += KI2[q] < ki ? 1u : 0u;
        K1[lo + rank] = k;
        V1[lo + rank] = V2[p];
    }
    __syncthreads();
    // 5. heads of equal-key runs fold their run left to right
    for (uint32_t e = tid; e < cnt; e += BK_THREADS) {
        const uint32_t k = K1[e];
        uint32_t keep = 0;
        if (e == 0 || K1[e - 1] != k) {
            double s = V1[e];
            for (uint32_t j = e + 1; j < cnt && K1[j] == k; ++j) s = __dadd_rn(s, V1[j]);
            V2[e] = s;
            keep = (s != 0.0) ? 1u : 0u;
        }
        I1[e] = keep;
    }
    rcnt[tid] = 0;
    __syncthreads();
    // 6. compact: thread t owns terms [t * BK_PER, (t + 1) * BK_PER)
    const uint32_t e0 = tid * BK_PER;
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < BK_PER; ++k)
        if (e0 + k < cnt) mine += I1[e0 + k];
    uint32_t nk;
    uint32_t o = bk_block_scan(mine, ws, nk);
#pragma unroll
    for (int k = 0; k < BK_PER; ++k) {
        const uint32_t e = e0 + k;
        if (e < cnt && I1[e]) {
            const uint32_t key = K1[e];
            tcol[s0 + o] = (int32_t)(key & colmask);
            tval[s0 + o] = V2[e];
            atomicAdd(&rcnt[key >> colbits], 1u);
            ++o;
        }
    }
    __syncthreads();
    if (tid < nr) rowcnt[row0 + tid] = rcnt[tid];
    if (tid == 0) bnnz[b] = nk;
}

// C2: survivors to their final place, row_ptr of the bucket's rows
static __global__ void __launch_bounds__(BK_THREADS)
bk_emit(const uint32_t* __restrict__ start, const uint32_t* __restrict__ bout, uint32_t nb, uint32_t rpb, uint32_t m,
        const int32_t* __restrict__ tcol, const double* __restrict__ tval, const uint32_t* __restrict__ rowcnt,
        int32_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx, double* __restrict__ vals) {
    __shared__ uint32_t ws[BK_THREADS / 32 + 1];
    const uint32_t b = blockIdx.x, tid = threadIdx.x;
    const uint32_t s0 = start[b], o0 = bout[b], nk = bout[b + 1] - o0;
    for (uint32_t k = tid; k < nk; k += BK_THREADS) {
        col_idx[o0 + k] = tcol[s0 + k];
        vals[o0 + k] = tval[s0 + k];
    }
    const uint32_t row0 = b * rpb;
    const uint32_t nr = min(rpb, m - row0);
    uint32_t tot;
    const uint32_t ex = bk_block_scan(tid < nr ? rowcnt[row0 + tid] : 0u, ws, tot);
    if (tid < nr) row_ptr[row0 + tid] = (int32_t)(o0 + ex);
    if (b == nb - 1 && tid == 0) row_ptr[m] = (int32_t)bout[nb];
}

struct BucketWorkspace {
    DevBuf<BkRec> rec;
    DevBuf<uint32_t> counts, start, cursor, bnnz, bout, stats, rowcnt;
    DevBuf<int32_t> tcol;
    DevBuf<double> tval;
};

inline int bk_bit_length(uint64_t v) {
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

// Returns false when the stream does not suit the bucket path (the outputs then hold nothing of value and the caller
// sorts); true: row_ptr, col_idx, vals are final and *nnz_out is nnz.  *bad_out != 0: a term lies outside the matrix.
// Synchronises the stream once, at the end.
inline bool assemble_bucketed(uint32_t T, const int32_t* d_row, const int32_t* d_col, const double* d_val, uint32_t m, uint32_t n,
                              int32_t* d_row_ptr, int32_t* d_col_idx, double* d_vals, uint32_t* nnz_out, int* d_bad,
                              int* bad_out, BucketWorkspace& w, cudaStream_t st) {
    const int colbits = bk_bit_length((uint64_t)n - 1);
    if (colbits > 31) return false;
    // rows per bucket: the mean bucket fills ~70 % of the capacity; the row inside a bucket and the column share 32 bits
    const double avg = (double)T / (double)m;
    uint64_t rpb64 = (uint64_t)std::max(1.0, std::floor(0.7 * BK_CAP / std::max(avg, 1e-9)));
    rpb64 = std::min<uint64_t>(rpb64, (uint64_t)BK_THREADS);
    rpb64 = std::min<uint64_t>(rpb64, 1ull << (32 - colbits));
    const uint32_t rpb = (uint32_t)rpb64;
    const uint64_t nb64 = ((uint64_t)m + rpb - 1) / rpb;
    if (nb64 > BK_MAX_BUCKETS) return false;
    const uint32_t nb = (uint32_t)nb64;
    auto grow = [](auto& b, size_t k) { if (b.n < k) b.alloc(k + k / 8); };
    grow(w.bnnz, nb); grow(w.bout, nb + 1); grow(w.stats, 4);
    grow(w.rec, T); grow(w.tcol, T); grow(w.tval, T); grow(w.rowcnt, (size_t)m + 1);
    ELP_CUDA(cudaMemsetAsync(w.stats.p, 0, 4 * sizeof(uint32_t), st));
    grow(w.counts, nb); grow(w.start, nb + 1); grow(w.cursor, nb);
    ELP_CUDA(cudaMemsetAsync(w.counts.p, 0, (size_t)nb * sizeof(uint32_t), st));
    const int grid = ceil_div((int64_t)T, 256);
    const size_t hist_bytes = (size_t)nb * sizeof(uint32_t);
    if (hist_bytes <= 96 * 1024) {
        // persistent grid, shared-memory histograms: two CTAs per SM when the histogram allows
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (200 * 1024) / (hist_bytes + 1024)));
        uint32_t G = (uint32_t)(kNumSMs * per_sm);
        G = std::max(1u, std::min(G, (T + 4095) / 4096));
        const uint32_t chunk = (T + G - 1) / G;
        static bool attr_set[16] = {};
        int dev = 0;
        ELP_CUDA(cudaGetDevice(&dev));
        if (!attr_set[dev & 15]) {
            ELP_CUDA(cudaFuncSetAttribute(bk_hist_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_set[dev & 15] = true;
        }
        ELP_LAUNCH(bk_hist_cta, G, BK_SPLIT_THREADS, hist_bytes, st, d_row, T, chunk, rpb, nb, m, w.counts.p, d_bad);
    } else {
        ELP_LAUNCH(bk_count, grid, 256, 0, st, d_row, T, rpb, m, w.counts.p, d_bad);
    }
    ELP_LAUNCH(bk_offsets, 1, 1024, 0, st, w.counts.p, nb, w.start.p, w.cursor.p, w.stats.p);
    ELP_LAUNCH(bk_scatter, grid, 256, 0, st, d_row, d_col, d_val, T, rpb, colbits, m, n, w.cursor.p, w.rec.p, d_bad);
    constexpr size_t smem = (size_t)BK_CAP * 32;
    {
        static bool attr_set[16] = {};
        int dev = 0;
        ELP_CUDA(cudaGetDevice(&dev));
        if (!attr_set[dev & 15]) {
            ELP_CUDA(cudaFuncSetAttribute(bk_sort_fold, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[dev & 15] = true;
        }
    }
    uint32_t* overflow = w.stats.p + 2;
    ELP_LAUNCH(bk_sort_fold, nb, BK_THREADS, smem, st, w.rec.p, w.start.p, rpb, colbits, m, w.tcol.p, w.tval.p, w.bnnz.p,
               w.rowcnt.p, overflow);
    ELP_LAUNCH(bk_offsets, 1, 1024, 0, st, w.bnnz.p, nb, w.bout.p, (uint32_t*)nullptr, w.stats.p);
    ELP_LAUNCH(bk_emit, nb, BK_THREADS, 0, st, w.start.p, w.bout.p, nb, rpb, m, w.tcol.p, w.tval.p, w.rowcnt.p, d_row_ptr,
               d_col_idx, d_vals);
    uint32_t stats[4] = {0, 0, 0, 0};
    ELP_CUDA(cudaMemcpyAsync(stats, w.stats.p, sizeof stats, cudaMemcpyDeviceToHost, st));
    ELP_CUDA(cudaMemcpyAsync(bad_out, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    ELP_CUDA(cudaStreamSynchronize(st));
    if (*bad_out) return false;                 // a term outside the matrix: the caller reports it
    if (stats[2] != 0) return false;            // a bucket held more than BK_CAP terms
    *nnz_out = stats[1];
    return true;
}

}  // namespace elp
