// simplex.cu — batched dense simplex: one LP per warp (one per CTA for the larger shapes), tableau resident in shared memory.
//
// Replaces `status <- solve(prob)` (/root/reference/R/class.R:276) for LPs whose dense tableau fits in
// one SM's shared memory, and is the engine of the additive batch entry point (BASELINE config 3:
// 200k LPs of 20x30).  Method: bounded-variable primal simplex on  A x + s = b  with a composite
// phase 1 (minimise the sum of infeasibilities), Dantzig pricing, a two-pass ratio test and Bland's
// rule after a run of degenerate pivots.  The CPU restatement is oracle/simplex_ref.c.
//
// Data movement: each LP's A (m*n doubles), b, c, lb, ub are read ONCE from HBM with 1-D TMA bulk
// copies (cp.async.bulk -> mbarrier complete_tx) straight into their shared-memory homes; results
// (status, objective, x) are written once.  Everything in between — pricing argmax, ratio-test
// argmin (warp shuffles), the rank-1 pivot update — runs out of shared memory in fp64.
// Algorithmic HBM bytes per LP: 8 (m n + m + 3 n) + m   in,   8 (n + 1) + 4   out.
// The serial pivot chain, not HBM, bounds this kernel; bench.py reports LPs/s, pivots/s and the HBM
// fraction side by side.
#include "common.cuh"
#include "tma.cuh"
#include "../../include/easylp_abi.h"
#include <algorithm>
#include <cmath>

namespace elp {

#define ST_BASIC 0
#define ST_LOWER 1
#define ST_UPPER 2
#define ST_FREE 3

constexpr double TOL_PRIMAL = 1e-9;
constexpr double TOL_DUAL = 1e-9;
constexpr double TOL_PIVOT = 1e-9;

__device__ __forceinline__ double ptol(double bound) { return TOL_PRIMAL * fmax(1.0, fabs(bound)); }

// ---- block-wide reductions (value, index) ----------------------------------------------------
struct ValIdx {
    double v;
    int i;
};
// "better" = larger v, ties -> smaller index.  Use negated values for argmin.
__device__ __forceinline__ ValIdx better(ValIdx a, ValIdx b) {
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
template <int THREADS>
__device__ __forceinline__ ValIdx block_argmax(ValIdx x, double* red_v, int* red_i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ValIdx y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        x = better(x, y);
    }
    if (THREADS == 32) return x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();   // protect scratch reuse
    if (lane == 0) { red_v[warp] = x.v; red_i[warp] = x.i; }
    __syncthreads();
    ValIdx r{red_v[0], red_i[0]};
#pragma unroll
    for (int w = 1; w < THREADS / 32; ++w) r = better(r, ValIdx{red_v[w], red_i[w]});
    return r;
}
template <int THREADS>
__device__ __forceinline__ double block_sum(double x, double* red_v) {
    x = warp_sum(x);
    if (THREADS == 32) return x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red_v[warp] = x;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) r += red_v[w];
    return r;
}
template <int THREADS>
__device__ __forceinline__ void block_sync() {
    if (THREADS == 32) __syncwarp();
    else __syncthreads();
}

struct SimplexSmemLayout {
    int m, n, N;
    size_t off_TA, off_TS, off_beta, off_lo, off_hi, off_cost, off_xn, off_cb, off_colq, off_rowr, off_redv;
    size_t off_basis, off_state, off_redi, off_bar, total;
    __host__ __device__ SimplexSmemLayout(int m_, int n_) : m(m_), n(n_), N(m_ + n_) {
        auto up2 = [](size_t v) { return (v + 1) & ~(size_t)1; };   // keep every double array 16B aligned
        size_t o = 0;
        off_TA = o;   o += up2((size_t)m * n);
        off_TS = o;   o += up2((size_t)m * m);
        off_beta = o; o += up2(m);
        off_lo = o;   o += up2(N);
        off_hi = o;   o += up2(N);
        off_cost = o; o += up2(N);
        off_xn = o;   o += up2(N);
        off_cb = o;   o += up2(m);
        off_colq = o; o += up2(m);
        off_rowr = o; o += up2(N);
        off_redv = o; o += 8;
        size_t bytes = o * sizeof(double);
        off_bar = bytes;   bytes += 16;
        off_basis = bytes; bytes += (size_t)((m + 3) & ~3) * sizeof(int);
        off_state = bytes; bytes += (size_t)((N + 3) & ~3) * sizeof(int);
        off_redi = bytes;  bytes += 8 * sizeof(int);
        total = bytes;
    }
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
simplex_batch_kernel(int64_t B, int m, int n, const double* __restrict__ Ag, const double* __restrict__ bg,
                     const double* __restrict__ cg, const double* __restrict__ lbg, const double* __restrict__ ubg,
                     const int8_t* __restrict__ senseg, int maximize, int max_pivots, int shared_model,
                     int32_t* __restrict__ status_out,
                     double* __restrict__ obj_out, double* __restrict__ x_out, double* __restrict__ y_out,
                     int32_t* __restrict__ pivots_out, int32_t* __restrict__ basis_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SimplexSmemLayout lay(m, n);
    const int N = lay.N;
    double* sd = reinterpret_cast<double*>(smem_raw);
    double* TA = sd + lay.off_TA;
    double* TS = sd + lay.off_TS;
    double* beta = sd + lay.off_beta;
    double* lo = sd + lay.off_lo;
    double* hi = sd + lay.off_hi;
    double* cost = sd + lay.off_cost;
    double* xn = sd + lay.off_xn;
    double* cb = sd + lay.off_cb;
    double* colq = sd + lay.off_colq;
    double* rowr = sd + lay.off_rowr;
    double* red_v = sd + lay.off_redv;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
    int* basis = reinterpret_cast<int*>(smem_raw + lay.off_basis);
    int* state = reinterpret_cast<int*>(smem_raw + lay.off_state);
    int* red_i = reinterpret_cast<int*>(smem_raw + lay.off_redi);
    const int tid = threadIdx.x;

    const int64_t lp = blockIdx.x;
    if (lp >= B) return;
    const int64_t mlp = shared_model ? 0 : lp;      // branch and bound: every node shares A, b, c, sense; bounds differ
    const double* A = Ag + mlp * (int64_t)m * n;
    const double* b = bg + mlp * (int64_t)m;
    const double* c = cg + mlp * (int64_t)n;
    const double* lb = lbg ? lbg + lp * (int64_t)n : nullptr;
    const double* ub = ubg ? ubg + lp * (int64_t)n : nullptr;
    const int8_t* sense = senseg ? senseg + mlp * (int64_t)m : nullptr;

    // ---- stage the LP: TMA bulk copies where alignment allows, cooperative loads otherwise ----------
    const size_t bytesA = (size_t)m * n * 8, bytesb = (size_t)m * 8, bytesn = (size_t)n * 8;
    const bool tA = tma_ok(A, bytesA), tb = tma_ok(b, bytesb), tc = tma_ok(c, bytesn);
    const bool tl = lb && tma_ok(lb, bytesn), tu = ub && tma_ok(ub, bytesn);
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    block_sync<THREADS>();
    if (tid == 0) {
        const uint32_t tx = (uint32_t)((tA ? bytesA : 0) + (tb ? bytesb : 0) + (tc ? bytesn : 0) + (tl ? bytesn : 0) +
                                       (tu ? bytesn : 0));
        mbar_expect_tx(bar, tx);
        if (tA) tma_load_1d(TA, A, (uint32_t)bytesA, bar);
        if (tb) tma_load_1d(beta, b, (uint32_t)bytesb, bar);
        if (tc) tma_load_1d(cost, c, (uint32_t)bytesn, bar);
        if (tl) tma_load_1d(lo, lb, (uint32_t)bytesn, bar);
        if (tu) tma_load_1d(hi, ub, (uint32_t)bytesn, bar);
    }
    // meanwhile: everything that does not depend on the copies
    for (int e = tid; e < m * m; e += THREADS) TS[e] = (e / m == e % m) ? 1.0 : 0.0;
    if (!tA) for (int e = tid; e < m * n; e += THREADS) TA[e] = A[e];
    if (!tb) for (int i = tid; i < m; i += THREADS) beta[i] = b[i];
    if (!tc) for (int j = tid; j < n; j += THREADS) cost[j] = c[j];
    if (!tl) for (int j = tid; j < n; j += THREADS) lo[j] = lb ? lb[j] : 0.0;
    if (!tu) for (int j = tid; j < n; j += THREADS) hi[j] = ub ? ub[j] : INFINITY;
    for (int i = tid; i < m; i += THREADS) {
        const int s = sense ? sense[i] : 0;
        lo[n + i] = (s == ELP_GE) ? -INFINITY : 0.0;
        hi[n + i] = (s == ELP_LE) ? INFINITY : 0.0;
        cost[n + i] = 0.0;
        xn[n + i] = 0.0;
        state[n + i] = ST_BASIC;
        basis[i] = n + i;
    }
    mbar_wait(bar, 0);
    block_sync<THREADS>();

    int bad = 0;
    for (int j = tid; j < n; j += THREADS) {
        if (maximize) cost[j] = -cost[j];
        const double l = lo[j], u = hi[j];
        if (l > u) bad = 1;
        if (isfinite(l)) { state[j] = ST_LOWER; xn[j] = l; }
        else if (isfinite(u)) { state[j] = ST_UPPER; xn[j] = u; }
        else { state[j] = ST_FREE; xn[j] = 0.0; }
    }
    bad = __syncthreads_or(bad);
    // beta = b - A xN
    for (int i = tid; i < m; i += THREADS) {
        double r = beta[i];
        for (int j = 0; j < n; ++j) r -= TA[i * n + j] * xn[j];
        beta[i] = r;
    }
    block_sync<THREADS>();

    int status = ELP_STATUS_TIMEOUT, pivots = 0, degenerate_run = 0, bland = 0;
    int qfinal = -1;
    if (max_pivots <= 0) max_pivots = 50 * (m + n) + 1000;
    if (bad) status = ELP_STATUS_INFEASIBLE;

    while (!bad) {
        // ---- phase detection ------------------------------------------------------------------
        double wpart = 0.0;
        for (int i = tid; i < m; i += THREADS) {
            const int k = basis[i];
            const double bi = beta[i];
            double g = 0.0;
            if (bi < lo[k] - ptol(lo[k])) { g = -1.0; wpart += lo[k] - bi; }
            else if (bi > hi[k] + ptol(hi[k])) { g = 1.0; wpart += bi - hi[k]; }
            cb[i] = g;
        }
        const double w = block_sum<THREADS>(wpart, red_v);
        const bool phase1 = w > 0.0;
        if (!phase1) for (int i = tid; i < m; i += THREADS) cb[i] = cost[basis[i]];
        block_sync<THREADS>();
        // ---- pricing: d_j = cost_j - cb' T[:,j]; Dantzig (largest |d_j|) or Bland (smallest j) ----
        ValIdx cand{0.0, -1};
        for (int j = tid; j < N; j += THREADS) {
            const int st = state[j];
            if (st == ST_BASIC || !(lo[j] < hi[j])) continue;
            double d = phase1 ? 0.0 : cost[j];
            if (j < n) { for (int i = 0; i < m; ++i) d -= cb[i] * TA[i * n + j]; }
            else { const int jj = j - n; for (int i = 0; i < m; ++i) d -= cb[i] * TS[i * m + jj]; }
            int dd = 0;
            if ((st == ST_LOWER || st == ST_FREE) && d < -TOL_DUAL) dd = 1;
            else if ((st == ST_UPPER || st == ST_FREE) && d > TOL_DUAL) dd = -1;
            if (!dd) continue;
            // encode the direction in the index: idx = 2*j + (dd<0)
            ValIdx me{bland ? -(double)j : fabs(d), 2 * j + (dd < 0 ? 1 : 0)};
            cand = better(cand, me);
        }
        cand = block_argmax<THREADS>(cand, red_v, red_i);
        if (cand.i < 0) { status = phase1 ? ELP_STATUS_INFEASIBLE : ELP_STATUS_OPTIMAL; break; }
        if (pivots >= max_pivots) { status = ELP_STATUS_TIMEOUT; break; }
        const int q = cand.i >> 1;
        const double dir = (cand.i & 1) ? -1.0 : 1.0;
        // ---- entering column ------------------------------------------------------------------
        for (int i = tid; i < m; i += THREADS) colq[i] = (q < n) ? TA[i * n + q] : TS[i * m + (q - n)];
        block_sync<THREADS>();
        // ---- ratio test, pass 1: the minimum step (warp-shuffle argmin) ----------------------------
        double tloc = INFINITY;
        for (int i = tid; i < m; i += THREADS) {
            const double a = dir * colq[i];
            if (fabs(a) <= TOL_PIVOT) continue;
            const int k = basis[i];
            const double bi = beta[i], l = lo[k], u = hi[k];
            double t = INFINITY;
            if (a > 0.0) {
                if (bi > u + ptol(u)) t = (bi - u) / a;
                else if (bi >= l - ptol(l)) { if (isfinite(l)) t = fmax(bi - l, 0.0) / a; }
            } else {
                if (bi < l - ptol(l)) t = (bi - l) / a;
                else if (bi <= u + ptol(u)) { if (isfinite(u)) t = fmin(bi - u, 0.0) / a; }
            }
            tloc = fmin(tloc, t);
        }
        ValIdx tm = block_argmax<THREADS>(ValIdx{-tloc, tid}, red_v, red_i);
        const double tmin = -tm.v;
        double tflip = (state[q] == ST_FREE) ? INFINITY : hi[q] - lo[q];
        if (tflip <= tmin) {
            if (!isfinite(tflip)) {
                if (phase1) { status = ELP_STATUS_NUMFAILURE; break; }
                status = ELP_STATUS_UNBOUNDED;
                qfinal = q;
                block_sync<THREADS>();
                if (tid == 0) xn[q] = dir > 0 ? INFINITY : -INFINITY;
                break;
            }
            for (int i = tid; i < m; i += THREADS) beta[i] -= dir * tflip * colq[i];
            block_sync<THREADS>();
            if (tid == 0) {
                if (dir > 0) { state[q] = ST_UPPER; xn[q] = hi[q]; }
                else { state[q] = ST_LOWER; xn[q] = lo[q]; }
            }
            ++pivots; degenerate_run = 0; bland = 0;
            block_sync<THREADS>();
            continue;
        }
        // ---- pass 2: among rows within a hair of tmin take the largest pivot (Bland: smallest basic id) ----
        const double window = tmin + 1e-12 * fmax(1.0, fabs(tmin));
        ValIdx rc{0.0, -1};
        for (int i = tid; i < m; i += THREADS) {
            const double a = dir * colq[i];
            if (fabs(a) <= TOL_PIVOT) continue;
            const int k = basis[i];
            const double bi = beta[i], l = lo[k], u = hi[k];
            double t = INFINITY; int up = 0;
            if (a > 0.0) {
                if (bi > u + ptol(u)) { t = (bi - u) / a; up = 1; }
                else if (bi >= l - ptol(l)) { if (isfinite(l)) { t = fmax(bi - l, 0.0) / a; up = 0; } }
            } else {
                if (bi < l - ptol(l)) { t = (bi - l) / a; up = 0; }
                else if (bi <= u + ptol(u)) { if (isfinite(u)) { t = fmin(bi - u, 0.0) / a; up = 1; } }
            }
            if (t > window) continue;
            ValIdx me{bland ? -(double)k : fabs(a), 2 * i + up};
            rc = better(rc, me);
        }
        rc = block_argmax<THREADS>(rc, red_v, red_i);
        if (rc.i < 0) { status = ELP_STATUS_NUMFAILURE; break; }
        const int r = rc.i >> 1;
        const int to_upper = rc.i & 1;
        // ---- step + basis change --------------------------------------------------------------
        const double t = tmin;
        const double piv = colq[r];
        for (int i = tid; i < m; i += THREADS) beta[i] -= dir * t * colq[i];
        for (int j = tid; j < N; j += THREADS) rowr[j] = ((j < n) ? TA[r * n + j] : TS[r * m + (j - n)]) / piv;
        block_sync<THREADS>();
        if (tid == 0) {
            const int kl = basis[r];
            if (to_upper) { state[kl] = ST_UPPER; xn[kl] = hi[kl]; }
            else { state[kl] = ST_LOWER; xn[kl] = lo[kl]; }
            beta[r] = xn[q] + dir * t;
            basis[r] = q;
            state[q] = ST_BASIC;
            rowr[q] = 1.0;
        }
        block_sync<THREADS>();
        // ---- rank-1 update of the tableau:  T -= colq * rowr  (row r := rowr) ----------------------
        for (int i = 0; i < m; ++i) {
            const double f = colq[i];
            if (i == r) {
                for (int j = tid; j < N; j += THREADS) { if (j < n) TA[i * n + j] = rowr[j]; else TS[i * m + (j - n)] = rowr[j]; }
            } else if (f != 0.0) {
                for (int j = tid; j < N; j += THREADS) {
                    double* e = (j < n) ? &TA[i * n + j] : &TS[i * m + (j - n)];
                    *e = (j == q) ? 0.0 : *e - f * rowr[j];      // the pivot column becomes a unit vector exactly
                }
            }
        }
        ++pivots;
        if (t <= 1e-12) { if (++degenerate_run > 30) bland = 1; }
        else { degenerate_run = 0; bland = 0; }
        block_sync<THREADS>();
    }
    block_sync<THREADS>();

    // ---- write back ---------------------------------------------------------------------------
    double* xo = x_out + lp * (int64_t)n;
    for (int j = tid; j < n; j += THREADS) if (state[j] != ST_BASIC) xo[j] = xn[j];
    for (int i = tid; i < m; i += THREADS) if (basis[i] < n) xo[basis[i]] = beta[i];
    double opart = 0.0;
    for (int j = tid; j < n; j += THREADS) {
        if (state[j] != ST_BASIC) opart += cost[j] * ((j == qfinal) ? 0.0 : xn[j]);
    }
    for (int i = tid; i < m; i += THREADS) if (basis[i] < n) opart += cost[basis[i]] * beta[i];
    double obj = block_sum<THREADS>(opart, red_v);
    if (status == ELP_STATUS_UNBOUNDED) obj = -INFINITY;
    if (tid == 0) {
        status_out[lp] = status;
        obj_out[lp] = maximize ? -obj : obj;
        if (pivots_out) pivots_out[lp] = pivots;
    }
    if (basis_out)      // final basis, one column id per row (structural j < n, slack of row i = n + i): sensitivity ranging
        for (int i = tid; i < m; i += THREADS) basis_out[lp * (int64_t)m + i] = basis[i];
    if (y_out) {
        double* yo = y_out + lp * (int64_t)m;
        for (int i = tid; i < m; i += THREADS) {
            double s = 0.0;
            for (int k = 0; k < m; ++k) s += cost[basis[k]] * TS[k * m + i];
            yo[i] = maximize ? -s : s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-per-LP kernel for small LPs (m <= 32, m + n <= 96) — the batched path of BASELINE config 3.
// One WARP solves one LP; a CTA is four independent warps.  The tableau T[MR][NS] (NS = 32*CPL + 2 doubles
// per row, MR = m rounded up to 4: both compile-time, so every tableau access is `base + immediate`) lives
// in the warp's slice of shared memory and is filled ONCE per LP by 1-D TMA bulk copies, one per row, issued
// by the lanes in parallel (cp.async.bulk -> mbarrier complete_tx); b, c, lb, ub arrive by coalesced loads.
// Lane l owns the tableau columns l, l+32, (l+64) and keeps their bounds / cost / non-basic value / state
// in registers; row-indexed data (basic values, basis heads, bounds of the basic variables, the entering
// column) sits beside the tableau.  A pivot is warp-synchronous end to end:
//   pricing     d_j = c_j - sum_i cb_i T[i][j]      one column per lane per pass, conflict-free row reads
//   argmax      three 32-bit hardware reductions (REDUX) on the bit pattern of |d_j| and on the index
//   ratio test  row i on lane i, REDUX min, second pass picks the largest pivot among the ties
//   update      T[i][j] -= colq[i] * rr[j]: MR fused multiply-adds per column straight on shared memory;
//               the dynamic pivot row is a plain address here (registers could not be indexed by it)
// There is no block-wide barrier, no address arithmetic in the loops and ~60 registers per thread, so many
// LPs are resident per SM.  Arithmetic order per entry equals simplex_batch_kernel / oracle/simplex_ref.c.
// ------------------------------------------------------------------------------------------------
constexpr int SW_WARPS = 4;

struct WarpSimplexLayout {      // per-warp shared memory, in bytes
    int ns, off_rows, total;
    __host__ __device__ WarpSimplexLayout(int MR, int CPL) {
        ns = 32 * CPL + 2;                                   // even (16-byte rows for TMA), not a multiple of 32
        off_rows = MR * ns * 8;
        // beta blo bhi bcost cb colq [MR each] | xs[96] | mbarrier | basis[MR]
        total = off_rows + 6 * MR * 8 + 96 * 8 + 16 + MR * 4;
        total = (total + 15) & ~15;
    }
};

struct WVI {
    double v;
    int i;
};
// Warp argmax of (v, i) with v >= 0 (or i < 0 = no candidate): larger v wins, ties go to the smaller i — the
// "better" of simplex_batch_kernel.  Non-negative doubles order like their bit patterns, so the maximum is two
// 32-bit hardware reductions (REDUX) on the high and low words and the tie-break a third on the index.
__device__ __forceinline__ WVI warp_best_nonneg(WVI x) {
    const unsigned long long bits = x.i < 0 ? 0ull : (unsigned long long)__double_as_longlong(x.v + 0.0);
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    const bool top = x.i >= 0 && hi == mhi && lo == mlo;
    const unsigned mi = __reduce_min_sync(0xffffffffu, top ? (unsigned)x.i : 0xffffffffu);
    WVI r;
    r.i = (mi == 0xffffffffu) ? -1 : (int)mi;
    r.v = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
    return r;
}
// Bland's rule: the candidate with the smallest key (a variable id), index i travels with it.
__device__ __forceinline__ int warp_best_bland(int key, int i) {
    const unsigned k = i < 0 ? 0xffffffffu : (unsigned)key;
    const unsigned mk = __reduce_min_sync(0xffffffffu, k);
    const unsigned mi = __reduce_min_sync(0xffffffffu, (i >= 0 && k == mk) ? (unsigned)i : 0xffffffffu);
    return mi == 0xffffffffu ? -1 : (int)mi;
}
// Warp minimum of non-negative doubles (+inf allowed)
__device__ __forceinline__ double warp_min_nonneg(double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v + 0.0);
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    return __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
}

template <int MR, int CPL>
__global__ void __launch_bounds__(SW_WARPS * 32)
simplex_warp_kernel(int64_t B, int m, int n, const double* __restrict__ Ag, const double* __restrict__ bg,
                    const double* __restrict__ cg, const double* __restrict__ lbg, const double* __restrict__ ubg,
                    const int8_t* __restrict__ senseg, int maximize, int max_pivots, int shared_model, int use_tma,
                    int32_t* __restrict__ status_out, double* __restrict__ obj_out, double* __restrict__ x_out,
                    double* __restrict__ y_out, int32_t* __restrict__ pivots_out, int32_t* __restrict__ basis_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NS = 32 * CPL + 2;
    const WarpSimplexLayout lay(MR, CPL);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + (size_t)warp * lay.total;
    double* Ts = reinterpret_cast<double*>(wbase);               // [MR][NS]
    double* Tl = Ts + lane;                                      // my columns: Tl[i * NS + 32 * c]
    double* rows = reinterpret_cast<double*>(wbase + lay.off_rows);
    double* beta = rows;
    double* blo = rows + MR;
    double* bhi = rows + 2 * MR;
    double* bcost = rows + 3 * MR;
    double* cb = rows + 4 * MR;
    double* colq = rows + 5 * MR;
    double* xs = rows + 6 * MR;                                  // [96] non-basic values by column (setup only)
    uint64_t* bar = reinterpret_cast<uint64_t*>(xs + 96);
    int* basis = reinterpret_cast<int*>(bar + 2);                // [MR]
    const int N = n + m;
    const int64_t gw = (int64_t)blockIdx.x * SW_WARPS + warp, nw = (int64_t)gridDim.x * SW_WARPS;
    if (max_pivots <= 0) max_pivots = 50 * (m + n) + 1000;

    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    // padding rows / columns stay zero for the whole kernel (updates of zero entries by zero factors)
    for (int e = lane; e < MR * NS; e += 32) Ts[e] = 0.0;
    __syncwarp();
    uint32_t phase = 0;

    for (int64_t lp = gw; lp < B; lp += nw) {
        const int64_t mlp = shared_model ? 0 : lp;  // branch and bound: every node shares A, b, c, sense; bounds differ
        const double* A = Ag + mlp * (int64_t)m * n;
        const int8_t* sense = senseg ? senseg + mlp * (int64_t)m : nullptr;
        // ---- tableau rows of A: one bulk copy per row, issued by the lanes in parallel ----------------
        if (use_tma && m > 0) {
            // generic-proxy writes to the tableau (previous LP) must be ordered before the async-proxy copies
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)m * (uint32_t)n * 8u);
            __syncwarp();
            for (int i = lane; i < m; i += 32) tma_load_1d(Ts + i * NS, A + (size_t)i * n, (uint32_t)n * 8u, bar);
        } else {
            for (int i = 0; i < m; ++i)
                for (int j = lane; j < n; j += 32) Ts[i * NS + j] = A[(size_t)i * n + j];
        }
        // ---- registers: scalars of my columns; slack block of the tableau -----------------------------
        double clo[CPL], chi[CPL], ccost[CPL], cxn[CPL];
        int cstate[CPL];
        int bad = 0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int col = lane + 32 * c;
            clo[c] = 0.0; chi[c] = 0.0; ccost[c] = 0.0; cxn[c] = 0.0; cstate[c] = ST_BASIC;
            if (col < n) {
                const double l = lbg ? lbg[lp * (int64_t)n + col] : 0.0, u = ubg ? ubg[lp * (int64_t)n + col] : INFINITY;
                const double cj = cg[mlp * (int64_t)n + col];
                clo[c] = l; chi[c] = u;
                ccost[c] = maximize ? -cj : cj;
                if (l > u) bad = 1;
                if (isfinite(l)) { cstate[c] = ST_LOWER; cxn[c] = l; }
                else if (isfinite(u)) { cstate[c] = ST_UPPER; cxn[c] = u; }
                else { cstate[c] = ST_FREE; cxn[c] = 0.0; }
                xs[col] = cxn[c];
            } else if (col < N) {
                const int i0 = col - n;
                const int s = sense ? sense[i0] : 0;
#pragma unroll
                for (int i = 0; i < MR; ++i) Tl[i * NS + 32 * c] = (i == i0) ? 1.0 : 0.0;
                clo[c] = (s == ELP_GE) ? -INFINITY : 0.0;
                chi[c] = (s == ELP_LE) ? INFINITY : 0.0;
                basis[i0] = col;
                blo[i0] = clo[c]; bhi[i0] = chi[c]; bcost[i0] = 0.0;
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        if (use_tma && m > 0) { mbar_wait(bar, phase); phase ^= 1u; }
        __syncwarp();
        // beta = b - A xN   (row i on lane i; sequential in j like the reference loop)
        if (lane < MR) {
            double r = 0.0;
            if (lane < m) {
                r = bg[mlp * (int64_t)m + lane];
                const double* Ti = Ts + lane * NS;
                for (int j = 0; j < n; ++j) r -= Ti[j] * xs[j];
            } else { blo[lane] = -INFINITY; bhi[lane] = INFINITY; bcost[lane] = 0.0; basis[lane] = -1; }
            beta[lane] = r;
            cb[lane] = 0.0; colq[lane] = 0.0;
        }
        __syncwarp();

        int status = ELP_STATUS_TIMEOUT, pivots = 0, degenerate_run = 0, bland = 0, qfinal = -1;
        if (bad) status = ELP_STATUS_INFEASIBLE;
        double d[CPL];                       // reduced costs of my columns (phase 2: carried across pivots)
        bool d_fresh_enough = false;
        int d_age = 0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) d[c] = 0.0;

        while (!bad) {
            // ---- phase detection: row i on lane i --------------------------------------------------
            double g = 0.0;
            if (lane < m) {
                const double bi = beta[lane], l = blo[lane], u = bhi[lane];
                if (bi < l - ptol(l)) g = -1.0;
                else if (bi > u + ptol(u)) g = 1.0;
            }
            // the sum of infeasibilities is positive iff some row is infeasible (every term is positive)
            const bool phase1 = __any_sync(0xffffffffu, g != 0.0);
            if (lane < m) cb[lane] = phase1 ? g : bcost[lane];
            __syncwarp();
            // ---- pricing on my columns.  In phase 2 the reduced costs d_j = c_j - cb' T_j are carried in registers
            // from pivot to pivot (d -= d_q * pivot row, the objective row of the textbook tableau) and recomputed
            // from the tableau every 32 pivots, on a phase change, and before optimality is declared. -----------
            WVI cand{0.0, -1};
            for (int attempt = 0; attempt < 2; ++attempt) {
                if (phase1 || !d_fresh_enough) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) d[c] = phase1 ? 0.0 : ccost[c];
#pragma unroll
                    for (int i = 0; i < MR; ++i) {
                        const double cbi = cb[i];
#pragma unroll
                        for (int c = 0; c < CPL; ++c) d[c] -= cbi * Tl[i * NS + 32 * c];
                    }
                    d_age = 0;
                }
                d_fresh_enough = !phase1;
                cand.v = 0.0; cand.i = -1;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int col = lane + 32 * c;
                    const int st = cstate[c];
                    if (col >= N || st == ST_BASIC || !(clo[c] < chi[c])) continue;
                    int dd = 0;
                    if ((st == ST_LOWER || st == ST_FREE) && d[c] < -TOL_DUAL) dd = 1;
                    else if ((st == ST_UPPER || st == ST_FREE) && d[c] > TOL_DUAL) dd = -1;
                    if (!dd) continue;
                    const double key = bland ? (double)col : fabs(d[c]);
                    const int id = 2 * col + (dd < 0 ? 1 : 0);
                    const bool take = cand.i < 0 || (bland ? key < cand.v : (key > cand.v || (key == cand.v && id < cand.i)));
                    if (take) { cand.v = key; cand.i = id; }
                }
                if (bland) cand.i = warp_best_bland((int)cand.v, cand.i);
                else cand = warp_best_nonneg(cand);
                if (cand.i >= 0 || d_age == 0) break;      // a candidate, or no candidate on freshly computed costs
                d_fresh_enough = false;                    // optimality is only declared on recomputed costs
            }
            if (cand.i < 0) { status = phase1 ? ELP_STATUS_INFEASIBLE : ELP_STATUS_OPTIMAL; break; }
            if (pivots >= max_pivots) { status = ELP_STATUS_TIMEOUT; break; }
            const int q = cand.i >> 1;
            const double dir = (cand.i & 1) ? -1.0 : 1.0;
            const int qlane = q & 31, qc = q >> 5;
            // ---- entering column: row i on lane i reads T[i][q]; its scalars come from the owner by shuffle ----
            double q_lo = 0.0, q_hi = 0.0, q_xn = 0.0, q_cost = 0.0;
            int q_state = 0;
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                if (c == qc) { q_lo = clo[c]; q_hi = chi[c]; q_xn = cxn[c]; q_cost = ccost[c]; q_state = cstate[c]; }
            q_lo = __shfl_sync(0xffffffffu, q_lo, qlane);
            q_hi = __shfl_sync(0xffffffffu, q_hi, qlane);
            q_xn = __shfl_sync(0xffffffffu, q_xn, qlane);
            q_cost = __shfl_sync(0xffffffffu, q_cost, qlane);
            q_state = __shfl_sync(0xffffffffu, q_state, qlane);
            double d_q = 0.0;
#pragma unroll
            for (int c = 0; c < CPL; ++c) if (c == qc) d_q = d[c];
            d_q = __shfl_sync(0xffffffffu, d_q, qlane);
            const double cq = (lane < MR) ? Ts[lane * NS + q] : 0.0;
            if (lane < MR) colq[lane] = cq;
            // ---- ratio test: row i on lane i ---------------------------------------------------------
            double tloc = INFINITY, a = 0.0;
            int up = 0, kbas = -1;
            if (lane < m) {
                a = dir * cq;
                if (fabs(a) > TOL_PIVOT) {
                    kbas = basis[lane];
                    const double bi = beta[lane], l = blo[lane], u = bhi[lane];
                    double num = INFINITY;                 // one division per row: t = num / a
                    if (a > 0.0) {
                        if (bi > u + ptol(u)) { num = bi - u; up = 1; }
                        else if (bi >= l - ptol(l) && isfinite(l)) { num = fmax(bi - l, 0.0); up = 0; }
                    } else {
                        if (bi < l - ptol(l)) { num = bi - l; up = 0; }
                        else if (bi <= u + ptol(u) && isfinite(u)) { num = fmin(bi - u, 0.0); up = 1; }
                    }
                    if (num != INFINITY) tloc = num / a;
                }
            }
            const double tmin = warp_min_nonneg(tloc);
            const double tflip = (q_state == ST_FREE) ? INFINITY : q_hi - q_lo;
            if (tflip <= tmin) {
                if (!isfinite(tflip)) {
                    if (phase1) { status = ELP_STATUS_NUMFAILURE; break; }
                    status = ELP_STATUS_UNBOUNDED;
                    qfinal = q;
#pragma unroll
                    for (int c = 0; c < CPL; ++c)
                        if (c == qc && lane == qlane) cxn[c] = dir > 0 ? INFINITY : -INFINITY;
                    break;
                }
                if (lane < m) beta[lane] -= dir * tflip * cq;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    if (c == qc && lane == qlane) {
                        if (dir > 0) { cstate[c] = ST_UPPER; cxn[c] = chi[c]; }
                        else { cstate[c] = ST_LOWER; cxn[c] = clo[c]; }
                    }
                }
                ++pivots; degenerate_run = 0; bland = 0;
                __syncwarp();
                continue;
            }
            // ---- pass 2: among rows within a hair of tmin take the largest pivot (Bland: smallest basic id) ---
            const double window = tmin + 1e-12 * fmax(1.0, fabs(tmin));
            WVI rc{0.0, -1};
            if (kbas >= 0 && tloc <= window) { rc.v = fabs(a); rc.i = 2 * lane + up; }
            if (bland) rc.i = warp_best_bland(kbas, rc.i);
            else rc = warp_best_nonneg(rc);
            if (rc.i < 0) { status = ELP_STATUS_NUMFAILURE; break; }
            const int r = rc.i >> 1;
            const int to_upper = rc.i & 1;
            const double t = tmin;
            __syncwarp();                                  // colq is complete
            const double piv = colq[r];
            const int kl = basis[r];                       // leaving variable (a column id)
            const double kl_bound = to_upper ? bhi[r] : blo[r];
            if (lane < m) beta[lane] -= dir * t * cq;
            // pivot row of my columns: the dynamic row is just an address in shared memory
            double rr[CPL];
            const double* Tr = Tl + r * NS;
#pragma unroll
            for (int c = 0; c < CPL; ++c) rr[c] = Tr[32 * c] / piv;
            __syncwarp();                                  // everyone has read row r before it is rewritten
            // bookkeeping by the owners
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = lane + 32 * c;
                if (col == kl) { cstate[c] = to_upper ? ST_UPPER : ST_LOWER; cxn[c] = kl_bound; }
                if (col == q) {
                    cstate[c] = ST_BASIC;
                    rr[c] = 1.0;
                    beta[r] = q_xn + dir * t;
                    basis[r] = q; blo[r] = q_lo; bhi[r] = q_hi; bcost[r] = q_cost;
                }
            }
            // ---- rank-1 update of my columns:  T[i] -= colq[i] * rr for i != r, row r := rr.  Column q needs no
            // special case: its rr is exactly 1 and its entries ARE colq, so f - f*1 = 0 exactly. ----------------
#pragma unroll
            for (int i = 0; i < MR; ++i) {
                const double f = colq[i];
#pragma unroll
                for (int c = 0; c < CPL; ++c) Tl[i * NS + 32 * c] -= f * rr[c];
            }
            {
                double* Trw = Tl + r * NS;
#pragma unroll
                for (int c = 0; c < CPL; ++c) Trw[32 * c] = rr[c];
            }
            // objective row: the entering column's reduced cost becomes exactly 0 (its rr is 1)
#pragma unroll
            for (int c = 0; c < CPL; ++c) d[c] -= d_q * rr[c];
            if (++d_age >= 32) d_fresh_enough = false;
            ++pivots;
            if (t <= 1e-12) { if (++degenerate_run > 30) bland = 1; }
            else { degenerate_run = 0; bland = 0; }
            __syncwarp();
        }
        __syncwarp();

        // ---- write back -----------------------------------------------------------------------------
        double* xo = x_out + lp * (int64_t)n;
        double opart = 0.0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int col = lane + 32 * c;
            if (col < n && cstate[c] != ST_BASIC) {
                xo[col] = cxn[c];
                opart += ccost[c] * ((col == qfinal) ? 0.0 : cxn[c]);
            }
        }
        if (lane < m) {
            const int k = basis[lane];
            if (k >= 0 && k < n) { xo[k] = beta[lane]; opart += bcost[lane] * beta[lane]; }
        }
        double obj = warp_sum(opart);
        if (status == ELP_STATUS_UNBOUNDED) obj = -INFINITY;
        if (lane == 0) {
            status_out[lp] = status;
            obj_out[lp] = maximize ? -obj : obj;
            if (pivots_out) pivots_out[lp] = pivots;
        }
        if (basis_out && lane < m) basis_out[lp * (int64_t)m + lane] = basis[lane];   // sensitivity ranging
        if (y_out) {
            // y_i = sum_k cost[basis[k]] * Tslack[k][i]: slack column n+i belongs to one lane
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = lane + 32 * c;
                if (col >= n && col < N) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < MR; ++k) s += bcost[k] * Tl[k * NS + 32 * c];
                    y_out[lp * (int64_t)m + (col - n)] = maximize ? -s : s;
                }
            }
        }
        __syncwarp();
        // leave the tableau clean for the next LP: structural columns are overwritten by the copies / loads,
        // padding never changes, the slack block is rewritten at setup — nothing to do.
    }
}

__global__ void k_densify(int m, int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                          const double* __restrict__ val, double* __restrict__ A) {
    const int i = blockIdx.x;
    if (i >= m) return;
    for (int k = ptr[i] + threadIdx.x; k < ptr[i + 1]; k += blockDim.x) atomicAdd(&A[(size_t)i * n + idx[k]], val[k]);
}

size_t simplex_smem_bytes(int m, int n) { return SimplexSmemLayout(m, n).total; }

static int env_flag(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int simplex_pick_threads(int m, int n) {
    const int N = m + n;
    if (const char* e = getenv("ELP_SIMPLEX_THREADS")) {
        const int t = atoi(e);
        if (t == 32 || t == 64 || t == 128) return t;
    }
    if (N <= 32) return 32;
    if (N <= 96) return 64;
    return 128;
}

// Warp-per-LP path: returns false when the shape does not fit the register tableau (caller falls back).
template <int MR, int CPL>
static void simplex_warp_launch_inst(int64_t B, int m, int n, const double* A, const double* b, const double* c,
                                     const double* lb, const double* ub, const int8_t* sense, int maximize, int max_pivots,
                                     int32_t* status, double* obj, double* x, double* y, int32_t* pivots, cudaStream_t st,
                                     int shared_model, int32_t* basis) {
    auto kern = simplex_warp_kernel<MR, CPL>;
    const WarpSimplexLayout lay(MR, CPL);
    const size_t smem = (size_t)SW_WARPS * lay.total;
    ELP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    ELP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SW_WARPS * 32, smem));
    occ = std::max(1, occ);
    auto al16 = [](const void* p) { return p == nullptr || (((uintptr_t)p) & 15) == 0; };
    const int use_tma = (n % 2 == 0) && al16(A) && env_flag("ELP_SIMPLEX_TMA", 1);   // 16-byte rows
    const int64_t want = (B + SW_WARPS - 1) / SW_WARPS;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)kNumSMs * occ));
    ELP_LAUNCH(kern, grid, SW_WARPS * 32, smem, st, B, m, n, A, b, c, lb, ub, sense, maximize, max_pivots, shared_model, use_tma,
               status, obj, x, y, pivots, basis);
}

static bool simplex_warp_launch(int64_t B, int m, int n, const double* A, const double* b, const double* c,
                                const double* lb, const double* ub, const int8_t* sense, int maximize, int max_pivots,
                                int32_t* status, double* obj, double* x, double* y, int32_t* pivots, cudaStream_t st,
                                int shared_model, int32_t* basis) {
    if (!env_flag("ELP_SIMPLEX_WARP", 1)) return false;
    const int N = m + n;
    if (m > 32 || N > 96) return false;
    const int mr = std::max(4, (m + 3) / 4 * 4), cpl = (N + 31) / 32;
    const size_t smem = (size_t)SW_WARPS * WarpSimplexLayout(mr, cpl).total;
    if (smem > 200 * 1024) return false;
#define ELP_SW(MR_, CPL_)                                                                                              \
    if (mr == MR_ && cpl == CPL_) {                                                                                    \
        simplex_warp_launch_inst<MR_, CPL_>(B, m, n, A, b, c, lb, ub, sense, maximize, max_pivots, status, obj, x, y, \
                                            pivots, st, shared_model, basis);                                          \
        return true;                                                                                                   \
    }
    ELP_SW(4, 1) ELP_SW(8, 1) ELP_SW(12, 1) ELP_SW(16, 1) ELP_SW(20, 1) ELP_SW(24, 1) ELP_SW(28, 1) ELP_SW(32, 1)
    ELP_SW(4, 2) ELP_SW(8, 2) ELP_SW(12, 2) ELP_SW(16, 2) ELP_SW(20, 2) ELP_SW(24, 2) ELP_SW(28, 2) ELP_SW(32, 2)
    ELP_SW(4, 3) ELP_SW(8, 3) ELP_SW(12, 3) ELP_SW(16, 3) ELP_SW(20, 3) ELP_SW(24, 3) ELP_SW(28, 3) ELP_SW(32, 3)
#undef ELP_SW
    return false;
}

// all pointers are device pointers
void simplex_batch_device(int64_t B, int m, int n, const double* A, const double* b, const double* c, const double* lb,
                          const double* ub, const int8_t* sense, int maximize, int max_pivots, int32_t* status,
                          double* obj, double* x, double* y, int32_t* pivots, cudaStream_t st, int shared_model,
                          int32_t* basis) {
    if (B <= 0) return;
    ELP_REQUIRE(n > 0 && m >= 0, "simplex: bad shape %d x %d", m, n);
    ELP_REQUIRE(B < 0x7fffffffll, "simplex: batch too large");
    if (simplex_warp_launch(B, m, n, A, b, c, lb, ub, sense, maximize, max_pivots, status, obj, x, y, pivots, st, shared_model, basis)) return;
    const size_t smem = simplex_smem_bytes(m, n);
    ELP_REQUIRE(smem <= 227 * 1024, "simplex: tableau of %d x %d needs %zu bytes of shared memory (max 227 KB)", m, n,
                smem);
    const int threads = simplex_pick_threads(m, n);
#define ELP_SIMPLEX_LAUNCH(T)                                                                                     \
    do {                                                                                                          \
        ELP_CUDA(cudaFuncSetAttribute(simplex_batch_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                      (int)smem));                                                                \
        ELP_LAUNCH((simplex_batch_kernel<T>), (unsigned)B, T, smem, st, B, m, n, A, b, c, lb, ub, sense, maximize, \
                   max_pivots, shared_model, status, obj, x, y, pivots, basis);                                   \
    } while (0)
    if (threads == 32) ELP_SIMPLEX_LAUNCH(32);
    else if (threads == 64) ELP_SIMPLEX_LAUNCH(64);
    else ELP_SIMPLEX_LAUNCH(128);
#undef ELP_SIMPLEX_LAUNCH
}

void densify_device(int m, int n, const int* ptr, const int* idx, const double* val, double* A, cudaStream_t st) {
    ELP_CUDA(cudaMemsetAsync(A, 0, (size_t)std::max(m, 1) * n * sizeof(double), st));
    if (m > 0) ELP_LAUNCH(k_densify, m, 64, 0, st, m, n, ptr, idx, val, A);
}

}  // namespace elp
