// assemble.cu — device-side assembly of expression terms into a canonical CSR matrix.
//
// Replaces the reference's dense constraint-matrix construction: `$con()` evaluating atoms and
// rbind-ing them (/root/reference/R/class.R:189-220, R/utils.R:95-106), where each dense entry was
// produced by adds folded left-to-right (R/methods.R:98-111 `horizontal_mat_sum`, R/methods.R:244-257
// `sum.lp_var` -> Reduce(`+`)).  Here the host emits one (row, col, val) term per non-zero
// contribution, in fold order; the device
//   1. builds keys row*n+col and a stable LSD radix sort groups equal (row,col) keeping emission order,
//   2. reduces each group strictly left-to-right (plain fp64 adds — no FMA, no re-association),
//   3. drops exact zeros (the reference's matrix simply has 0 there),
//   4. prefix-sums the keep flags and scatters (col, val), then fills row_ptr.
// Result: bit-identical to the != 0 entries of the reference's dense `constraint$mat`.
//
// Roofline: HBM-bound streaming passes.  Algorithmic bytes: 16*T + 12*nnz + 4*(m+1).
#include "common.cuh"
#include <cstdlib>
#include "primitives.cuh"
#include "bucket_sort.cuh"
#include "../../include/easylp_abi.h"

namespace elp {

// bad[0]: a term lies outside the matrix; bad[1]: the stream is NOT already in (row, col) order.  `$con()` emits its
// rows one after the other and most bodies walk their columns in ascending order, so the common stream is sorted as it
// arrives: the host then skips the radix sort altogether (the fold only needs equal keys to be adjacent, in emission order).
__global__ void asm_make_keys(const int32_t* __restrict__ row, const int32_t* __restrict__ col, uint32_t T,
                              uint32_t m, uint32_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ perm,
                              int* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint32_t r = (uint32_t)row[i], c = (uint32_t)col[i];
    const bool oob = (r >= m || c >= n);
    if (oob) atomicExch(bad, 1);
    const uint64_t k = oob ? 0ull : (uint64_t)r * n + c;   // out-of-range terms are reported, never dereferenced
    keys[i] = k;
    perm[i] = i;
    if (i + 1 < T) {
        const uint32_t r1 = (uint32_t)row[i + 1], c1 = (uint32_t)col[i + 1];
        const uint64_t k1 = (r1 >= m || c1 >= n) ? 0ull : (uint64_t)r1 * n + c1;
        if (k1 < k && bad[1] == 0) atomicExch(bad + 1, 1);
    }
}

// 4096 adjacent pairs spread over the stream: one inversion proves the stream unordered before anything is written for it
__global__ void asm_sample_order(const int32_t* __restrict__ row, const int32_t* __restrict__ col, uint32_t T, int* __restrict__ flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (T < 2) return;
    const uint32_t i = (uint32_t)(((uint64_t)t * (T - 1)) >> 12);
    const uint32_t r0 = (uint32_t)row[i], r1 = (uint32_t)row[i + 1];
    if (r1 < r0 || (r1 == r0 && (uint32_t)col[i + 1] < (uint32_t)col[i])) *flag = 1;
}

// ---- ordered streams: (row, col) never decreases, so no keys and no permutation are needed ------------------------------
// `$con()` emits its rows one after the other and most bodies walk their columns upwards.  Two passes straight over the
// stream: (1) bounds, order and, per tile of 1 024 terms, how many folded entries survive and the row of the last one;
// (2) after a scan of the tile counts the same fold again, survivors written to their final place together with the
// row pointers (an entry whose predecessor among the survivors sits in another row starts the rows in between).
// 32 + 12 bytes per term instead of the 80 of the key / fold / scan / compact pipeline (0.74 ms for the 21 M terms of
// config 4).  A thread owns four consecutive terms (vector loads; one term per thread ran both passes at 2.1 TB/s).
constexpr int AO_V = 4, AO_THREADS = 256, AO_TILE = AO_V * AO_THREADS;
struct AoTerms {
    int32_t r[AO_V], c[AO_V];
    double v[AO_V];
    int32_t pr, pc, nr, nc;          // neighbours of the four: the term before the first, the term behind the last
    uint32_t cnt;                    // terms of the four that exist (the stream may end inside)
};
// loads the thread's terms and their neighbours; out-of-stream slots repeat sentinels that never compare equal
__device__ __forceinline__ void ao_load(const int32_t* __restrict__ row, const int32_t* __restrict__ col, const double* __restrict__ val,
                                        uint32_t T, uint64_t i0, int lane, bool vec, AoTerms& t) {
    t.cnt = i0 >= T ? 0u : (uint32_t)min((uint64_t)AO_V, (uint64_t)T - i0);
    if (t.cnt == AO_V && vec) {               // vec: the three arrays start on 16-byte boundaries
        const int4 rr = *reinterpret_cast<const int4*>(row + i0), cc = *reinterpret_cast<const int4*>(col + i0);
        const double2 v0 = *reinterpret_cast<const double2*>(val + i0), v1 = *reinterpret_cast<const double2*>(val + i0 + 2);
        t.r[0] = rr.x; t.r[1] = rr.y; t.r[2] = rr.z; t.r[3] = rr.w;
        t.c[0] = cc.x; t.c[1] = cc.y; t.c[2] = cc.z; t.c[3] = cc.w;
        t.v[0] = v0.x; t.v[1] = v0.y; t.v[2] = v1.x; t.v[3] = v1.y;
    } else {
#pragma unroll
        for (int k = 0; k < AO_V; ++k) {
            const bool in = (uint32_t)k < t.cnt;
            t.r[k] = in ? row[i0 + k] : -2; t.c[k] = in ? col[i0 + k] : -2; t.v[k] = in ? val[i0 + k] : 0.0;
        }
    }
    // neighbours: the lane below / above holds them, the warp's ends read them from memory
    int32_t pr = __shfl_up_sync(0xffffffffu, t.r[AO_V - 1], 1), pc = __shfl_up_sync(0xffffffffu, t.c[AO_V - 1], 1);
    int32_t nr = __shfl_down_sync(0xffffffffu, t.r[0], 1), nc = __shfl_down_sync(0xffffffffu, t.c[0], 1);
    if (lane == 0) {
        const bool has = i0 > 0 && i0 <= T;
        pr = has ? row[i0 - 1] : -1; pc = has ? col[i0 - 1] : -1;
    }
    if (lane == 31) {
        const bool has = i0 + AO_V < T;
        nr = has ? row[i0 + AO_V] : -3; nc = has ? col[i0 + AO_V] : -3;
    }
    if (t.cnt < AO_V) { nr = -3; nc = -3; }       // the stream ends inside my four (the slots behind carry -2)
    t.pr = pr; t.pc = pc; t.nr = nr; t.nc = nc;
}
// folds the runs that START inside the thread's four: keep[k] / s[k] for their heads
__device__ __forceinline__ void ao_fold(const int32_t* __restrict__ row, const int32_t* __restrict__ col, const double* __restrict__ val,
                                        uint32_t T, uint64_t i0, const AoTerms& t, bool* keep, double* s) {
#pragma unroll
    for (int k = 0; k < AO_V; ++k) {
        keep[k] = false; s[k] = 0.0;
        if ((uint32_t)k >= t.cnt) continue;
        const int32_t r = t.r[k], c = t.c[k];
        const bool head = k == 0 ? !(t.pr == r && t.pc == c) : !(t.r[k - 1] == r && t.c[k - 1] == c);
        if (!head) continue;
        double acc = t.v[k];
        bool open = true;                         // the run is still going
#pragma unroll
        for (int j = k + 1; j < AO_V; ++j) {
            open = open && (uint32_t)j < t.cnt && t.r[j] == r && t.c[j] == c;
            if (open) acc = __dadd_rn(acc, t.v[j]);
        }
        if (open && t.nr == r && t.nc == c)       // it runs on behind my four: straight from the stream, left to right
            for (uint64_t j = i0 + AO_V; j < T && row[j] == r && col[j] == c; ++j) acc = __dadd_rn(acc, val[j]);
        s[k] = acc;
        keep[k] = acc != 0.0;
    }
}
__global__ void __launch_bounds__(AO_THREADS)
asm_ordered_count(const int32_t* __restrict__ row, const int32_t* __restrict__ col, const double* __restrict__ val, uint32_t T,
                  uint32_t m, uint32_t n, uint32_t* __restrict__ tile_cnt, int32_t* __restrict__ tile_last, int* __restrict__ bad,
                  bool vec) {
    __shared__ int skip;                 // an inversion is already known (the sample launched just before, or another
    __shared__ uint32_t wcnt[AO_THREADS / 32];
    __shared__ int wlast[AO_THREADS / 32];
    if (threadIdx.x == 0) skip = bad[1]; // block of this pass): nothing this block computes will be used
    __syncthreads();
    if (skip) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t i0 = ((uint64_t)blockIdx.x * AO_THREADS + threadIdx.x) * AO_V;
    AoTerms t;
    ao_load(row, col, val, T, i0, lane, vec, t);
    bool oob = false, inv = false;
#pragma unroll
    for (int k = 0; k < AO_V; ++k) {
        if ((uint32_t)k >= t.cnt) continue;
        oob |= (uint32_t)t.r[k] >= m || (uint32_t)t.c[k] >= n;
        const bool has_next = k + 1 < AO_V ? (uint32_t)(k + 1) < t.cnt : t.nr != -3;
        const uint32_t r1 = (uint32_t)(k + 1 < AO_V ? t.r[k + 1] : t.nr), c1 = (uint32_t)(k + 1 < AO_V ? t.c[k + 1] : t.nc);
        inv |= has_next && (r1 < (uint32_t)t.r[k] || (r1 == (uint32_t)t.r[k] && c1 < (uint32_t)t.c[k]));
    }
    if (oob) bad[0] = 1;
    if (inv) bad[1] = 1;
    bool keep[AO_V];
    double s[AO_V];
    ao_fold(row, col, val, T, i0, t, keep, s);
    uint32_t mine = 0;
    int last = -1;
#pragma unroll
    for (int k = 0; k < AO_V; ++k)
        if (keep[k]) { ++mine; last = t.r[k]; }
    mine = __reduce_add_sync(0xffffffffu, mine);
    last = __reduce_max_sync(0xffffffffu, last);          // rows never decrease: the largest is the last
    if (lane == 0) { wcnt[warp] = mine; wlast[warp] = last; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        int l = -1;
        for (int w = 0; w < AO_THREADS / 32; ++w) { c += wcnt[w]; l = max(l, wlast[w]); }
        tile_cnt[blockIdx.x] = c;
        tile_last[blockIdx.x] = l;
    }
}
__global__ void __launch_bounds__(AO_THREADS)
asm_ordered_emit(const int32_t* __restrict__ row, const int32_t* __restrict__ col, const double* __restrict__ val, uint32_t T,
                 const uint32_t* __restrict__ tile_off, const int32_t* __restrict__ tile_last, int32_t* __restrict__ col_idx,
                 double* __restrict__ vals, int32_t* __restrict__ row_ptr, uint32_t* __restrict__ nnz_out,
                 int32_t* __restrict__ last_row_out, const int* __restrict__ bad, bool vec) {
    __shared__ uint32_t wsum[AO_THREADS / 32];
    __shared__ int wlast[AO_THREADS / 32];
    __shared__ int base_last, near, skip;
    if (threadIdx.x == 0) { skip = bad[0] | bad[1]; base_last = -1; near = 0x7fffffff; }
    __syncthreads();
    if (skip) return;                    // not ordered: the host takes another path; a term outside the matrix: the host
                                         // reports it (its row must never index row_ptr)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t i0 = ((uint64_t)blockIdx.x * AO_THREADS + threadIdx.x) * AO_V;
    AoTerms t;
    ao_load(row, col, val, T, i0, lane, vec, t);
    bool keep[AO_V];
    double s[AO_V];
    ao_fold(row, col, val, T, i0, t, keep, s);
    uint32_t mine = 0;
    int last = -1;
#pragma unroll
    for (int k = 0; k < AO_V; ++k)
        if (keep[k]) { ++mine; last = t.r[k]; }
    // inside the warp: exclusive sum of the counts, exclusive max of the last surviving rows
    uint32_t inc = mine;
    int lmax = last;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, inc, o);
        const int b = __shfl_up_sync(0xffffffffu, lmax, o);
        if (lane >= o) { inc += a; lmax = max(lmax, b); }
    }
    int pr = __shfl_up_sync(0xffffffffu, lmax, 1);
    if (lane == 0) pr = -1;
    if (lane == 31) { wsum[warp] = inc; wlast[warp] = lmax; }
    // the last surviving row before this tile: normally the previous tile's; the block looks back 256 tiles per round
    for (int64_t top = (int64_t)blockIdx.x - 1; top >= 0; top -= AO_THREADS) {
        const int64_t tt = top - threadIdx.x;
        if (tt >= 0 && tile_last[tt] >= 0) atomicMin(&near, (int)threadIdx.x);
        __syncthreads();
        const int k = near;
        __syncthreads();                         // everybody has read `near` before the next round may lower it
        if (k != 0x7fffffff) {
            if (threadIdx.x == 0) base_last = tile_last[top - k];
            break;                               // uniform: every thread read the same `near`
        }
    }
    __syncthreads();
    uint32_t p = tile_off[blockIdx.x] + inc - mine;
    int before = base_last;
    for (int w = 0; w < warp; ++w) {
        p += wsum[w];
        before = max(before, wlast[w]);
    }
    pr = max(pr, before);
#pragma unroll
    for (int k = 0; k < AO_V; ++k) {
        if (!keep[k]) continue;
        col_idx[p] = t.c[k]; vals[p] = s[k];
        for (int32_t rr = pr + 1; rr <= t.r[k]; ++rr) row_ptr[rr] = (int32_t)p;
        pr = t.r[k];
        ++p;
    }
    if (i0 < T && i0 + AO_V >= T) {              // the thread that holds the last term
        *nnz_out = p;
        *last_row_out = pr;                      // asm_row_ptr_tail closes the rows behind it
    }
}
__global__ void asm_row_ptr_tail(const int32_t* __restrict__ last_row, const uint32_t* __restrict__ nnz, uint32_t m, int32_t* __restrict__ row_ptr,
                                 const int* __restrict__ bad) {
    if (bad[0] | bad[1]) return;         // the emit pass did not run
    const int64_t rr = (int64_t)*last_row + 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (rr >= 0 && rr <= (int64_t)m) row_ptr[rr] = (int32_t)*nnz;
}

// One thread per sorted slot; the head of every equal-key run folds the run left-to-right.
__global__ void asm_segment_fold(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                                 const double* __restrict__ val, uint32_t T, double* __restrict__ sums,
                                 uint32_t* __restrict__ keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint64_t k = keys[i];
    if (i > 0 && keys[i - 1] == k) { keep[i] = 0; return; }
    double s = val[perm[i]];
    for (uint32_t j = i + 1; j < T && keys[j] == k; ++j) s = __dadd_rn(s, val[perm[j]]);
    sums[i] = s;
    keep[i] = (s != 0.0) ? 1u : 0u;
}

// Two-level ordered fold (lowered assembly): inside a (row, col) run the terms arrive group-major (stable sort of a
// group-major stream).  Each group is summed left to right, multiplied by its post-fold multipliers, dropped if it is an
// exact zero (the host's canonical partial sums carry no zeros), and the group results are added left to right.
__device__ __forceinline__ double asm_finish_group(double gs, uint32_t g, uint32_t row, const elp_fold_group* __restrict__ groups,
                                                   const double* __restrict__ dtab) {
    if (g == 0) return gs;
    const elp_fold_group* G = groups + g;          // read in place: a local copy of the struct would live on the stack
    const int n_mul = G->n_mul;
    for (int k = 0; k < n_mul; ++k)
        gs = __dmul_rn(gs, dtab[G->mul_tab[k] + (G->mul_per_row[k] ? (int64_t)(row - (uint32_t)G->row0) : 0)]);
    return gs;
}
__global__ void asm_segment_fold_grouped(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                                         const double* __restrict__ val, const int32_t* __restrict__ grp,
                                         const elp_fold_group* __restrict__ groups, const double* __restrict__ dtab,
                                         uint32_t T, uint32_t n, double* __restrict__ sums, uint32_t* __restrict__ keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint64_t k = keys[i];
    if (i > 0 && keys[i - 1] == k) { keep[i] = 0; return; }
    const uint32_t row = (uint32_t)(k / n);
    uint32_t p = perm[i];
    uint32_t g = (uint32_t)grp[p];
    double gs = val[p], acc = 0.0;
    bool have = false;
    for (uint32_t j = i + 1; j < T && keys[j] == k; ++j) {
        p = perm[j];
        const uint32_t gj = (uint32_t)grp[p];
        if (gj == g) { gs = __dadd_rn(gs, val[p]); continue; }
        gs = asm_finish_group(gs, g, row, groups, dtab);
        if (gs != 0.0) { acc = have ? __dadd_rn(acc, gs) : gs; have = true; }
        g = gj; gs = val[p];
    }
    gs = asm_finish_group(gs, g, row, groups, dtab);
    if (gs != 0.0) { acc = have ? __dadd_rn(acc, gs) : gs; have = true; }
    sums[i] = acc;
    keep[i] = (have && acc != 0.0) ? 1u : 0u;
}

// One thread per term of a family: decode the position in the loop nest (loops slowest first), accumulate the row,
// the column offset and the coefficient index, write the term at its place in the stream.
__global__ void asm_expand_family(const elp_term_family f, const int32_t* __restrict__ itab, const double* __restrict__ dtab,
                                  int32_t* __restrict__ row, int32_t* __restrict__ col, double* __restrict__ val,
                                  int32_t* __restrict__ grp) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= f.count) return;
    int64_t rem = idx, ci = f.coef_tab;
    int32_t r = f.row0, c = f.col0;
    for (int l = f.n_loops - 1; l >= 0; --l) {
        const int32_t e = f.extent[l];
        const int32_t p = (int32_t)(rem % e);
        rem /= e;
        r += (f.row_tab[l] >= 0) ? itab[f.row_tab[l] + p] : f.row_stride[l] * p;
        if (f.col_tab[l] >= 0) c += itab[f.col_tab[l] + p];
        ci += f.coef_stride[l] * p;
    }
    const int64_t pos = f.out_offset + idx * f.out_stride;
    row[pos] = r; col[pos] = c; val[pos] = dtab[ci]; grp[pos] = f.group;
}

__global__ void asm_compact(const uint64_t* __restrict__ keys, const double* __restrict__ sums,
                            const uint32_t* __restrict__ keep_flag, const uint32_t* __restrict__ pos, uint32_t T,
                            uint32_t n, int32_t* __restrict__ col_idx, double* __restrict__ vals,
                            int32_t* __restrict__ rows_c, uint32_t* __restrict__ nnz_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint32_t p = pos[i];
    if (keep_flag[i]) {
        const uint64_t k = keys[i];
        const uint32_t r = (uint32_t)(k / n);
        col_idx[p] = (int32_t)(k - (uint64_t)r * n);
        vals[p] = sums[i];
        rows_c[p] = (int32_t)r;
    }
    if (i == T - 1) *nnz_out = p + keep_flag[i];
}

// thread i in [0, nnz]: fills row_ptr for the rows that start at compacted position i
__global__ void asm_row_ptr(const int32_t* __restrict__ rows_c, const uint32_t* __restrict__ nnz_ptr, uint32_t m,
                            int32_t* __restrict__ row_ptr, uint32_t T) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nnz = *nnz_ptr;
    if (i > nnz || i > T) return;
    const int32_t prev = (i == 0) ? -1 : rows_c[i - 1];
    const int32_t cur = (i == nnz) ? (int32_t)m : rows_c[i];
    for (int32_t r = prev + 1; r <= cur; ++r) row_ptr[r] = (int32_t)i;
}

static bool env_flag(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e ? atoi(e) : dflt) != 0;
}

__global__ void asm_empty_row_ptr(int32_t* row_ptr, uint32_t m) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= m) row_ptr[i] = 0;
}

// Grow-only device workspace of the assembly (per host thread; re-created when the thread switches device).
struct AsmWorkspace {
    int device = -1;
    DevBuf<uint64_t> keys;
    DevBuf<uint32_t> perm, keep, pos, nnz_d;
    DevBuf<double> sums;
    DevBuf<int32_t> rows_c;
    DevBuf<int> bad;
    RadixSortWorkspace sort;
    BucketWorkspace bucket;
    // staging of the host entry point (elp_assemble_csr)
    DevBuf<int32_t> in_row, in_col, out_ptr, out_col;
    DevBuf<double> in_val, out_val;
    // lowered assembly: fold groups of the stream, descriptor tables
    DevBuf<int32_t> in_grp;
    DevBuf<int32_t> itab;
    DevBuf<double> dtab;
    DevBuf<elp_fold_group> groups;
    void release() { *this = AsmWorkspace{}; }
    void ensure(size_t T) {
        int dev = 0;
        ELP_CUDA(cudaGetDevice(&dev));
        if (dev != device) { release(); device = dev; }
        auto grow = [](auto& b, size_t n) { if (b.n < n) b.alloc(n + n / 8); };
        grow(keys, T); grow(perm, T); grow(keep, T); grow(pos, T); grow(sums, T); grow(rows_c, T);
        grow(nnz_d, 1); grow(bad, 2);
    }
    void ensure_io(size_t T, size_t m) {
        ensure(T);
        auto grow = [](auto& b, size_t n) { if (b.n < n) b.alloc(n + n / 8); };
        grow(in_row, T); grow(in_col, T); grow(in_val, T); grow(out_col, T); grow(out_val, T); grow(out_ptr, m + 1);
    }
};
AsmWorkspace& asm_workspace() {
    static thread_local AsmWorkspace w;
    return w;
}
void asm_workspace_release() { asm_workspace().release(); }
struct AsmIo {
    int32_t *row, *col, *out_ptr, *out_col;
    double *val, *out_val;
};
struct AsmLowered {
    int32_t* grp;
    int32_t* itab;
    double* dtab;
    elp_fold_group* groups;
};
AsmIo asm_io_buffers(size_t T, size_t m) {
    AsmWorkspace& w = asm_workspace();
    w.ensure_io(T, m);
    return AsmIo{w.in_row.p, w.in_col.p, w.out_ptr.p, w.out_col.p, w.in_val.p, w.out_val.p};
}
AsmLowered asm_lowered_buffers(size_t T, size_t n_itab, size_t n_dtab, size_t n_groups) {
    AsmWorkspace& w = asm_workspace();
    auto grow = [](auto& b, size_t n) { if (b.n < n) b.alloc(n + n / 8); };
    grow(w.in_grp, std::max<size_t>(T, 1)); grow(w.itab, std::max<size_t>(n_itab, 1));
    grow(w.dtab, std::max<size_t>(n_dtab, 1)); grow(w.groups, std::max<size_t>(n_groups, 1));
    return AsmLowered{w.in_grp.p, w.itab.p, w.dtab.p, w.groups.p};
}
void asm_expand_families(int n_families, const elp_term_family* fam, const int32_t* d_itab, const double* d_dtab,
                         int64_t stream_offset, int32_t* d_row, int32_t* d_col, double* d_val, int32_t* d_grp,
                         cudaStream_t st) {
    for (int i = 0; i < n_families; ++i) {
        elp_term_family f = fam[i];
        if (f.count <= 0) continue;
        f.out_offset += stream_offset;
        ELP_LAUNCH(asm_expand_family, ceil_div(f.count, 256), 256, 0, st, f, d_itab, d_dtab, d_row, d_col, d_val, d_grp);
    }
}

// Device-resident assembly: inputs/outputs are device pointers.  d_col_idx/d_vals need room for T.
// Returns nnz (synchronises the stream once to read it back).
int64_t assemble_csr_device(int64_t T, const int32_t* d_row, const int32_t* d_col, const double* d_val, int32_t m,
                            int32_t n, int32_t* d_row_ptr, int32_t* d_col_idx, double* d_vals, cudaStream_t st,
                            const int32_t* d_grp, const elp_fold_group* d_groups, const double* d_dtab) {
    ELP_REQUIRE(m >= 0 && n >= 0, "assemble: negative shape");
    if (T == 0 || m == 0) {
        ELP_LAUNCH(asm_empty_row_ptr, ceil_div((int64_t)m + 1, 256), 256, 0, st, d_row_ptr, (uint32_t)m);
        ELP_CUDA(cudaStreamSynchronize(st));
        ELP_REQUIRE(T == 0, "assemble: %lld terms but the matrix has no rows", (long long)T);
        return 0;
    }
    ELP_REQUIRE(T < 0xffffffffll, "assemble: too many terms");
    ELP_REQUIRE(n > 0, "assemble: terms but no columns");
    const uint32_t Tu = (uint32_t)T;
    WallTimer dbg_t;
    const bool dbg = getenv("ELP_ASM_DEBUG") != nullptr;
    auto mark = [&](const char* what) {
        if (!dbg) return;
        cudaStreamSynchronize(st);
        fprintf(stderr, "[assemble] %-14s %9.3f ms\n", what, dbg_t.ms());
    };
    // Temporaries come from a grow-only per-thread workspace: cudaMalloc/cudaFree of ~50 B per term cost tens of
    // milliseconds per call, far more than the kernels (3.2 ms for 21 M terms).
    AsmWorkspace& w = asm_workspace();
    w.ensure(T);
    DevBuf<uint64_t>& keys = w.keys;
    DevBuf<uint32_t>&perm = w.perm, &keep = w.keep, &pos = w.pos, &nnz_d = w.nnz_d;
    DevBuf<double>& sums = w.sums;
    DevBuf<int32_t>& rows_c = w.rows_c;
    DevBuf<int>& bad = w.bad;
    ELP_CUDA(cudaMemsetAsync(bad.p, 0, 2 * sizeof(int), st));
    RadixSortWorkspace& ws = w.sort;
    const int grid = ceil_div(T, 256);
    mark("alloc");
    const bool always_sort = getenv("ELP_ASM_ALWAYS_SORT") != nullptr;
    // Unordered plain streams: one split into row buckets + a shared-memory sort-and-fold per bucket (bucket_sort.cuh)
    // instead of six global radix passes.  Grouped (lowered) folds and streams with a bucket that does not fit keep the sort.
    const bool may_bucket = !always_sort && !d_grp && T >= 2048 && env_flag("ELP_ASM_BUCKETED", 1);
    auto try_buckets = [&](int64_t* nnz_out) {
        uint32_t nnz_b = 0;
        int bad_b = 0;
        const bool ok = assemble_bucketed(Tu, d_row, d_col, d_val, (uint32_t)m, (uint32_t)n, d_row_ptr, d_col_idx, d_vals, &nnz_b,
                                          bad.p, &bad_b, w.bucket, st);
        ELP_REQUIRE(!bad_b, "assemble: a term has row/col outside [0,%d) x [0,%d)", m, n);
        mark(ok ? "bucketed" : "bucket path refused");
        *nnz_out = (int64_t)nnz_b;
        return ok;
    };
    bool tried = false;
    if (!always_sort && !d_grp) {
        int inv = 0;
        // a sample of adjacent pairs first: an inversion there makes the full pass return at once (no synchronisation
        // in between: the pass reads the flag on the device)
        if (T >= 2048) ELP_LAUNCH(asm_sample_order, 16, 256, 0, st, d_row, d_col, Tu, bad.p + 1);
        {   // bounds, order and the tile counts; scan; fold + emit — launched back to back, the later ones return at once when
            // an inversion is flagged.  One synchronisation, at the end.
            const int tiles = ceil_div(T, AO_TILE);
            int32_t* tile_last = reinterpret_cast<int32_t*>(keep.p);          // one word per tile (the workspace holds T)
            const bool vec = (((uintptr_t)d_row | (uintptr_t)d_col | (uintptr_t)d_val) & 15u) == 0;
            ELP_LAUNCH(asm_ordered_count, tiles, AO_THREADS, 0, st, d_row, d_col, d_val, Tu, (uint32_t)m, (uint32_t)n, pos.p, tile_last,
                       bad.p, vec);
            exclusive_scan_u32(pos.p, (size_t)tiles, ws.scan, st);
            ELP_LAUNCH(asm_ordered_emit, tiles, AO_THREADS, 0, st, d_row, d_col, d_val, Tu, pos.p, tile_last, d_col_idx, d_vals, d_row_ptr,
                       nnz_d.p, rows_c.p, bad.p, vec);
            // rows behind the last surviving entry (all rows when nothing survived)
            ELP_LAUNCH(asm_row_ptr_tail, ceil_div((int64_t)m + 1, 256), 256, 0, st, rows_c.p, nnz_d.p, (uint32_t)m, d_row_ptr, bad.p);
            int flags[2] = {0, 0};
            uint32_t nnz_o = 0;
            ELP_CUDA(cudaMemcpyAsync(flags, bad.p, sizeof flags, cudaMemcpyDeviceToHost, st));
            ELP_CUDA(cudaMemcpyAsync(&nnz_o, nnz_d.p, sizeof nnz_o, cudaMemcpyDeviceToHost, st));
            ELP_CUDA(cudaStreamSynchronize(st));
            inv = flags[1];
            if (!inv) {
                ELP_REQUIRE(!flags[0], "assemble: a term has row/col outside [0,%d) x [0,%d)", m, n);
                mark("ordered");
                return (int64_t)nnz_o;
            }
            mark("not ordered");
        }
        if (inv && may_bucket) {
            int64_t nnz_b = 0;
            if (try_buckets(&nnz_b)) return nnz_b;
            tried = true;
        }
    }
    ELP_LAUNCH(asm_make_keys, grid, 256, 0, st, d_row, d_col, Tu, (uint32_t)m, (uint32_t)n, keys.p, perm.p, bad.p);
    const int nbits = bit_length_u64((uint64_t)m * (uint64_t)n - 1);
    mark("keys");
    int flags[2] = {0, 1};                           // [0] a term outside the matrix, [1] the stream is not in (row, col) order
    if (!always_sort) {                              // 8 bytes back and one synchronisation buy the whole sort when the stream is ordered
        ELP_CUDA(cudaMemcpyAsync(flags, bad.p, sizeof flags, cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        ELP_REQUIRE(!flags[0], "assemble: a term has row/col outside [0,%d) x [0,%d)", m, n);
    }
    const int unsorted = flags[1];
    if (unsorted && may_bucket && !tried) {          // the sample saw no inversion, the full pass did
        int64_t nnz_b = 0;
        if (try_buckets(&nnz_b)) return nnz_b;
    }
    if (unsorted) radix_sort_pairs(keys.p, perm.p, T, nbits, ws, st);
    mark(unsorted ? "sort" : "sort skipped");
    if (d_grp)
        ELP_LAUNCH(asm_segment_fold_grouped, grid, 256, 0, st, keys.p, perm.p, d_val, d_grp, d_groups, d_dtab, Tu, (uint32_t)n,
                   sums.p, keep.p);
    else
        ELP_LAUNCH(asm_segment_fold, grid, 256, 0, st, keys.p, perm.p, d_val, Tu, sums.p, keep.p);
    ELP_CUDA(cudaMemcpyAsync(pos.p, keep.p, T * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    mark("fold");
    exclusive_scan_u32(pos.p, T, ws.scan, st);
    mark("scan");
    ELP_LAUNCH(asm_compact, grid, 256, 0, st, keys.p, sums.p, keep.p, pos.p, Tu, (uint32_t)n, d_col_idx, d_vals,
               rows_c.p, nnz_d.p);
    ELP_LAUNCH(asm_row_ptr, ceil_div(T + 1, 256), 256, 0, st, rows_c.p, nnz_d.p, (uint32_t)m, d_row_ptr, Tu);
    uint32_t nnz = 0;
    int bad_h = 0;
    ELP_CUDA(cudaMemcpyAsync(&nnz, nnz_d.p, sizeof nnz, cudaMemcpyDeviceToHost, st));
    ELP_CUDA(cudaMemcpyAsync(&bad_h, bad.p, sizeof bad_h, cudaMemcpyDeviceToHost, st));
    ELP_CUDA(cudaStreamSynchronize(st));
    mark("compact+ptr");
    ELP_REQUIRE(!bad_h, "assemble: a term has row/col outside [0,%d) x [0,%d)", m, n);
    return (int64_t)nnz;
}

}  // namespace elp
