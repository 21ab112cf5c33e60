// pdlp.cu — PDLP-style first-order LP solver for large sparse LPs on one or several B200s.
//
// Replaces `status <- solve(prob)` (/root/reference/R/class.R:276; lp_solve's simplex) for LPs whose
// dense tableau does not fit in shared memory.  Problem form (after mapping the reference's rows,
// R/class.R:271-274):   min c'x   s.t.  lc <= A x <= uc ,  l <= x <= u.
//
// Algorithm: restarted reflected Halpern PDHG (r2HPDHG; Lu & Yang 2024, the cuPDLPx scheme):
//   z' = T(z):  x' = proj_[l,u](x - tau (c - A'y)),   y' = prox(y - sigma A(2x' - x))
//   z+ = (k+1)/(k+2) (2 z' - z) + 1/(k+2) z_anchor          (k = iterations since the last restart)
// with Ruiz + Pock-Chambolle diagonal scaling, constant step eta = 0.998/||A||_2 (tau = eta/w,
// sigma = eta*w), primal weight w re-balanced at restarts, fixed-point-error restarts and PDLP's
// relative KKT termination test.  The CPU restatement of this loop is oracle/pdlp_ref.py.
//
// Kernels per iteration (single GPU) — exactly two, both HBM-bound SpMVs with fused epilogues (spmv.cuh):
//   K1  CSC (transposed) SpMV  g = A'y  + projection + reflection + Halpern combine  -> x, xbar
//   K2  CSR SpMV  A xbar       + dual prox + reflection + Halpern combine            -> y
// Matrix stream, row pointers and the epilogue operands arrive in shared memory by TMA bulk copies;
// the gathered vector (x: 8n bytes, y: 8m bytes) lives in the 126 MB L2.
// Algorithmic bytes per iteration: 24 nnz + 4 (m+n+2) + 56 n + 40 m.
//
// Multi-GPU: rank g owns a row block of A (for K2, its slice of y) AND a column block of A over all
// rows (for K1, its slice of x).  Nothing is replicated; per iteration the x-bar blocks and the y blocks are
// all-gathered over NVLink (8 (n + m) bytes received per rank) and the scalar residuals of a check travel in one
// 16-double allreduce.  See `struct Pdlp`.
#include "common.cuh"
#include "primitives.cuh"
#include "comm.cuh"
#include "tma.cuh"
#include "spmv.cuh"
#include "../../include/easylp_abi.h"
#include <cmath>
#include <algorithm>

namespace elp {

struct PdlpParams {   // device-resident; the host rewrites it between iteration chunks
    double tau, sigma;
    int k_base;       // iterations since restart at the start of the chunk
    int pad;
};

constexpr int SPMV_THREADS = 256;             // block size of the setup-only helper kernels

// ---- epilogues of the SpMV (spmv.cuh): in(i) names the operand vectors the producer stages next to the
// matrix stream, preload() picks this row's operands out of the stage, apply() runs after the row sum ------
struct StoreEpi {
    static constexpr int NIN = 0;
    static constexpr bool SCATTER = false;
    double* scat = nullptr;
    double* out;
    struct Pre {};
    __device__ __forceinline__ const double* in(int) const { return nullptr; }
    __device__ __forceinline__ Pre preload(const double*, int, int) const { return Pre{}; }
    __device__ __forceinline__ Pre preload_global(int) const { return Pre{}; }
    __device__ __forceinline__ double apply(int r, double s, const Pre&, const L2Hints&) const { out[r] = s; return 0.0; }
};

inline StoreEpi store_epi(double* out) { StoreEpi e; e.out = out; return e; }

// Where an epilogue publishes the block it produces: its own copy of the gathered vector, or — when the ranks have
// mapped each other's buffers (CUDA IPC over NVLink) — the copy of EVERY rank, so that the all-gather is done by the
// stores of the kernel itself, overlapped with the rest of the tile walk.
struct PeerOut {
    double* p[8];
    int n;          // 0: single destination (the local pointer of the epilogue)
    // mask[j] bit r: rank r gathers entry j of this block (its matrix block references it).  Entries nobody else reads
    // stay at home: on structured LPs (multi-commodity flow: a row block touches the columns of its own commodities)
    // most of the exchange disappears.  nullptr = every rank gets every entry.
    const unsigned char* mask;
};

// primal half of T(z) + reflection + Halpern combine.  g = (A'y)_j
template <bool CHECK>
struct PrimalEpi {
    static constexpr int NIN = CHECK ? 4 : 5;
    static constexpr bool SCATTER = false;
    double* scat;
    const double* __restrict__ c;
    const double* __restrict__ l;
    const double* __restrict__ u;
    const double* __restrict__ x0;
    double* __restrict__ x;
    double* __restrict__ xbar;
    double* __restrict__ xp;
    const PdlpParams* __restrict__ P;
    int it;
    PeerOut peers;
    struct Pre { double x, c, l, u, x0; };
    __device__ __forceinline__ const double* in(int i) const {
        return i == 0 ? x : i == 1 ? c : i == 2 ? l : i == 3 ? u : x0;
    }
    __device__ __forceinline__ Pre preload(const double* s, int rt, int g) const {
        Pre p;
        p.x = s[g]; p.c = s[rt + g]; p.l = s[2 * rt + g]; p.u = s[3 * rt + g];
        p.x0 = CHECK ? 0.0 : s[(CHECK ? 0 : 4) * rt + g];
        return p;
    }
    __device__ __forceinline__ Pre preload_global(int j) const {
        Pre p;
        p.x = x[j]; p.c = c[j]; p.l = l[j]; p.u = u[j];
        p.x0 = CHECK ? 0.0 : x0[j];
        return p;
    }
    __device__ __forceinline__ double apply(int j, double g, const Pre& p, const L2Hints& h) const {
        const double tau = P->tau;
        const double xpj = fmin(fmax(p.x - tau * (p.c - g), p.l), p.u);
        const double xb = 2.0 * xpj - p.x;
        if (!CHECK && peers.n > 0) {
            const unsigned mk = peers.mask ? peers.mask[j] : 0xffu;
#pragma unroll 8
            for (int r = 0; r < peers.n; ++r)
                if ((mk >> r) & 1u) peers.p[r][j] = xb;
        } else {
            st_out(xbar + j, xb, h, true);
        }
        if (CHECK) {
            xp[j] = xpj;
        } else {
            const double k = (double)(P->k_base + it);
            const double w = (k + 1.0) / (k + 2.0);
            st_out(x + j, w * xb + (1.0 - w) * p.x0, h, false);
        }
        return 0.0;
    }
};

// dual half.  ax = (A xbar)_i.  SCAT: the kernel then adds val[k] * y_new_i into scat[idx[k]] over the entries of row i,
// i.e. it leaves g = A'y_new behind for the next (gather-free) primal update.
template <bool CHECK, bool SCAT = false>
struct DualEpi {
    static constexpr int NIN = CHECK ? 3 : 4;
    static constexpr bool SCATTER = SCAT;
    double* scat;
    const double* __restrict__ lc;
    const double* __restrict__ uc;
    const double* __restrict__ y0;
    double* __restrict__ y;
    double* __restrict__ yp;
    double* __restrict__ axbar;
    const PdlpParams* __restrict__ P;
    int it;
    PeerOut peers;
    struct Pre { double y, lc, uc, y0; };
    __device__ __forceinline__ const double* in(int i) const { return i == 0 ? y : i == 1 ? lc : i == 2 ? uc : y0; }
    __device__ __forceinline__ Pre preload(const double* s, int rt, int g) const {
        Pre p;
        p.y = s[g]; p.lc = s[rt + g]; p.uc = s[2 * rt + g];
        p.y0 = CHECK ? 0.0 : s[(CHECK ? 0 : 3) * rt + g];
        return p;
    }
    __device__ __forceinline__ Pre preload_global(int i) const {
        Pre p;
        p.y = y[i]; p.lc = lc[i]; p.uc = uc[i];
        p.y0 = CHECK ? 0.0 : y0[i];
        return p;
    }
    __device__ __forceinline__ double apply(int i, double ax, const Pre& p, const L2Hints& h) const {
        const double sigma = P->sigma;
        const double v = p.y - sigma * ax;
        const double lo = v + sigma * p.lc;     // -inf when the row has no lower bound
        const double hi = v + sigma * p.uc;     // +inf when the row has no upper bound
        const double ypi = lo > 0.0 ? lo : (hi < 0.0 ? hi : 0.0);
        if (CHECK) {
            yp[i] = ypi;
            axbar[i] = ax;
        } else {
            const double k = (double)(P->k_base + it);
            const double w = (k + 1.0) / (k + 2.0);
            const double yn = w * (2.0 * ypi - p.y) + (1.0 - w) * p.y0;
            if (peers.n > 0) {
                const unsigned mk = peers.mask ? peers.mask[i] : 0xffu;
#pragma unroll 8
                for (int r = 0; r < peers.n; ++r)
                    if ((mk >> r) & 1u) peers.p[r][i] = yn;
            } else {
                st_out(y + i, yn, h, true);
            }
            return yn;
        }
        return 0.0;
    }
};

// Gather-free primal update of the scatter formulation: g = A'y was accumulated by the previous dual kernel.
// Reads g, x, c, l, u, x0; writes x-bar, x and clears g for the next accumulation.  Two columns per thread (16-byte accesses).
__global__ void __launch_bounds__(256)
k_primal_from_g(int n, double* __restrict__ g, const double* __restrict__ c, const double* __restrict__ l,
                const double* __restrict__ u, const double* __restrict__ x0, double* __restrict__ x,
                double* __restrict__ xbar, const PdlpParams* __restrict__ P, int it) {
    const double tau = P->tau;
    const double k = (double)(P->k_base + it);
    const double w = (k + 1.0) / (k + 2.0);
    const int npair = n >> 1;
    for (int q = blockIdx.x * 256 + threadIdx.x; q < npair; q += gridDim.x * 256) {
        const double2 gv = reinterpret_cast<const double2*>(g)[q];
        const double2 xv = reinterpret_cast<const double2*>(x)[q];
        const double2 cv = reinterpret_cast<const double2*>(c)[q];
        const double2 lv = reinterpret_cast<const double2*>(l)[q];
        const double2 uv = reinterpret_cast<const double2*>(u)[q];
        const double2 av = reinterpret_cast<const double2*>(x0)[q];
        const double p0 = fmin(fmax(xv.x - tau * (cv.x - gv.x), lv.x), uv.x);
        const double p1 = fmin(fmax(xv.y - tau * (cv.y - gv.y), lv.y), uv.y);
        const double b0 = 2.0 * p0 - xv.x, b1 = 2.0 * p1 - xv.y;
        reinterpret_cast<double2*>(xbar)[q] = make_double2(b0, b1);
        reinterpret_cast<double2*>(x)[q] = make_double2(w * b0 + (1.0 - w) * av.x, w * b1 + (1.0 - w) * av.y);
        reinterpret_cast<double2*>(g)[q] = make_double2(0.0, 0.0);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int j = n - 1;
        const double xj = x[j];
        const double pj = fmin(fmax(xj - tau * (c[j] - g[j]), l[j]), u[j]);
        const double bj = 2.0 * pj - xj;
        xbar[j] = bj;
        x[j] = w * bj + (1.0 - w) * x0[j];
        g[j] = 0.0;
    }
}

// lanes per row for the setup-only helper kernels (row statistics, value scaling, row expansion)
inline int pick_helper_lanes(int64_t nnz, int64_t nrows) {
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    int L = 1;
    while (L < 32 && (double)(2 * L) <= avg * 0.75 + 1.0) L *= 2;   // avg 10 -> 8, avg 5 -> 4, avg 2 -> 1..2
    return L;
}

// ---- row statistics for the scaling: out[r] = s_self[r] * reduce_k |val[k]| * s_other[idx[k]] ------
template <int L, int MODE /*0 max, 1 sum*/>
__global__ void __launch_bounds__(SPMV_THREADS)
rowstat_kernel(int nrows, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
               const double* __restrict__ s_self, const double* __restrict__ s_other, double* __restrict__ out) {
    const int row = (int)(((int64_t)blockIdx.x * SPMV_THREADS + threadIdx.x) / L);
    const int lane = threadIdx.x & (L - 1);
    int start = 0, end = 0;
    if (row < nrows) { start = ptr[row]; end = ptr[row + 1]; }
    double s = 0.0;
    for (int k = start + lane; k < end; k += L) {
        const double a = fabs(val[k]) * s_other[idx[k]];
        s = MODE == 0 ? fmax(s, a) : s + a;
    }
    s = MODE == 0 ? group_max<L>(s) : group_sum<L>(s);
    if (lane == 0 && row < nrows) out[row] = s * s_self[row];
}

template <int L>
__global__ void __launch_bounds__(SPMV_THREADS)
scale_vals_kernel(int nrows, const int* __restrict__ ptr, const int* __restrict__ idx, double* __restrict__ val,
                  const double* __restrict__ s_self, const double* __restrict__ s_other) {
    const int row = (int)(((int64_t)blockIdx.x * SPMV_THREADS + threadIdx.x) / L);
    const int lane = threadIdx.x & (L - 1);
    if (row >= nrows) return;
    const double sr = s_self[row];
    for (int k = ptr[row] + lane; k < ptr[row + 1]; k += L) val[k] *= sr * s_other[idx[k]];
}

template <int L>
__global__ void __launch_bounds__(SPMV_THREADS)
expand_rows_kernel(int nrows, const int* __restrict__ ptr, int* __restrict__ row_of) {
    const int row = (int)(((int64_t)blockIdx.x * SPMV_THREADS + threadIdx.x) / L);
    const int lane = threadIdx.x & (L - 1);
    if (row >= nrows) return;
    for (int k = ptr[row] + lane; k < ptr[row + 1]; k += L) row_of[k] = row;
}

// ---- small elementwise kernels ----------------------------------------------------------------
__global__ void k_ruiz_update(int n, double* __restrict__ d, const double* __restrict__ stat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double s = stat[i]; if (s > 0.0) d[i] *= rsqrt(s); }
}
__global__ void k_fill(int n, double* __restrict__ d, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = v;
}
__global__ void k_mul(int n, double* __restrict__ d, const double* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] *= s[i];
}
__global__ void k_div(int n, double* __restrict__ d, const double* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] /= s[i];
}
__global__ void k_scale_scalar(int n, double* __restrict__ d, double a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] *= a;
}
__global__ void k_row_bounds(int m, const int8_t* __restrict__ sense, const double* __restrict__ rhs,
                             double* __restrict__ lc, double* __restrict__ uc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int s = sense[i];
    const double r = rhs[i];
    lc[i] = (s == ELP_LE) ? -INFINITY : r;
    uc[i] = (s == ELP_GE) ? INFINITY : r;
}
__global__ void k_init_x(int n, double* __restrict__ x, const double* __restrict__ l, const double* __restrict__ u) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) x[j] = fmin(fmax(0.0, l[j]), u[j]);
}
__global__ void k_pseudo_random(int n, double* __restrict__ v, uint32_t seed) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t h = (uint32_t)j * 2654435761u + seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    v[j] = (double)h / 4294967296.0 - 0.5;
}
// after a check without restart: finish the Halpern step from the stored candidate
__global__ void k_halpern_finish(int n, double* __restrict__ x, const double* __restrict__ xbar,
                                 const double* __restrict__ x0, int m, double* __restrict__ y,
                                 const double* __restrict__ yp, const double* __restrict__ y0, double w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = w * xbar[i] + (1.0 - w) * x0[i];
    if (i < m) y[i] = w * (2.0 * yp[i] - y[i]) + (1.0 - w) * y0[i];
}
// restart: iterate and anchor both jump to the candidate T(z)
__global__ void k_restart(int n, double* __restrict__ x, double* __restrict__ x0, const double* __restrict__ xp, int m,
                          double* __restrict__ y, double* __restrict__ y0, const double* __restrict__ yp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double v = xp[i]; x[i] = v; x0[i] = v; }
    if (i < m) { const double v = yp[i]; y[i] = v; y0[i] = v; }
}
__global__ void k_unscale(int n, const double* __restrict__ v, const double* __restrict__ s, double sign,
                          double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sign * v[i] * s[i];
}

// ---- reductions: per-block partials in fixed order, then one block folds them --------------------
constexpr int RED_THREADS = 256;
constexpr int RED_BLOCKS = kNumSMs * 4;
constexpr int NACC = 8;

__device__ __forceinline__ void block_reduce_store(double (&acc)[NACC], double* __restrict__ partials) {
    __shared__ double sm[RED_THREADS / 32][NACC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < RED_THREADS / 32; ++w) s += sm[w][threadIdx.x];
        partials[blockIdx.x * NACC + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(RED_THREADS)
k_final_reduce(const double* __restrict__ partials, int nblocks, double* __restrict__ out) {
    __shared__ double sm[RED_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < NACC; ++k) {
        double s = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += RED_THREADS) s += partials[b * NACC + k];
        s = warp_sum(s);
        if (lane == 0) sm[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < RED_THREADS / 32; ++w) t += sm[w];
            out[k] = t;
        }
        __syncthreads();
    }
}

// row-side check sums.  acc: 0 pres^2 (unscaled), 1 dy.(A dx), 2 |dy|^2, 3 |yp-y0|^2, 4 dual obj (rows),
//                            5 |b|^2 helper unused, 6 ray: |yp - y0|_inf proxy unused, 7 unused
__global__ void __launch_bounds__(RED_THREADS)
k_check_rows(int m, const double* __restrict__ axp, const double* __restrict__ axbar, const double* __restrict__ y,
             const double* __restrict__ yp, const double* __restrict__ y0, const double* __restrict__ lc,
             const double* __restrict__ uc, const double* __restrict__ dr, double* __restrict__ partials) {
    double acc[NACC] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < m; i += gridDim.x * RED_THREADS) {
        const double a = axp[i], lo = lc[i], hi = uc[i];
        const double viol = (a - fmin(fmax(a, lo), hi)) / dr[i];
        const double ypi = yp[i];
        const double dy = ypi - y[i];
        const double d0 = ypi - y0[i];
        acc[0] += viol * viol;
        acc[1] += dy * (axbar[i] - a);
        acc[2] += dy * dy;
        acc[3] += d0 * d0;
        acc[4] += ypi > 0.0 ? ypi * lo : (ypi < 0.0 ? ypi * hi : 0.0);
    }
    block_reduce_store(acc, partials);
}

// column-side check sums.  acc: 0 dres^2 (unscaled), 1 |dx|^2, 2 |xp-x0|^2, 3 c'xp, 4 dual obj (bounds)
__global__ void __launch_bounds__(RED_THREADS)
k_check_cols(int n, const double* __restrict__ g, const double* __restrict__ c, const double* __restrict__ l,
             const double* __restrict__ u, const double* __restrict__ x, const double* __restrict__ xp,
             const double* __restrict__ x0, const double* __restrict__ dc, double* __restrict__ partials) {
    double acc[NACC] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = blockIdx.x * RED_THREADS + threadIdx.x; j < n; j += gridDim.x * RED_THREADS) {
        const double r = c[j] - g[j];
        const double xpj = xp[j], lj = l[j], uj = u[j];
        const bool at_lo = isfinite(lj) && xpj <= lj;
        const bool at_hi = isfinite(uj) && xpj >= uj;
        const double rpos = fmax(r, 0.0), rneg = fmin(r, 0.0);
        const double res = ((at_lo ? 0.0 : rpos) + (at_hi ? 0.0 : rneg)) / dc[j];
        const double dx = xpj - x[j];
        const double d0 = xpj - x0[j];
        acc[0] += res * res;
        acc[1] += dx * dx;
        acc[2] += d0 * d0;
        acc[3] += c[j] * xpj;
        acc[4] += (at_lo ? rpos * lj : 0.0) + (at_hi ? rneg * uj : 0.0);
    }
    block_reduce_store(acc, partials);
}

// generic: acc0 = sum a^2, acc1 = sum a*b (b may alias a), for norms / dot products
__global__ void __launch_bounds__(RED_THREADS)
k_dot2(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ partials) {
    double acc[NACC] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += gridDim.x * RED_THREADS) {
        const double ai = a[i];
        acc[0] += ai * ai;
        acc[1] += ai * b[i];
    }
    block_reduce_store(acc, partials);
}

// sum of squares of the finite row-bound magnitudes (PDLP's ||b||) -> acc0 ; acc1 = sum c^2 style helper
__global__ void __launch_bounds__(RED_THREADS)
k_bound_norm(int m, const double* __restrict__ lc, const double* __restrict__ uc, double* __restrict__ partials) {
    double acc[NACC] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < m; i += gridDim.x * RED_THREADS) {
        const double lo = lc[i], hi = uc[i];
        const double a = fmax(isfinite(lo) ? fabs(lo) : 0.0, isfinite(hi) ? fabs(hi) : 0.0);
        acc[0] += a * a;
    }
    block_reduce_store(acc, partials);
}

// certificate sums on the displacement ray (dx = xp - x0 over columns)
//  cols: 0 c'dx, 1 bound violation^2 of the ray, 2 max|dx| (as sum of squares -> use norm), 3 |dx|^2
__global__ void __launch_bounds__(RED_THREADS)
k_ray_cols(int n, const double* __restrict__ xp, const double* __restrict__ x0, const double* __restrict__ c,
           const double* __restrict__ l, const double* __restrict__ u, const double* __restrict__ gray,
           const double* __restrict__ dc, double* __restrict__ partials) {
    // gray = A'(yp - y0) (scaled).  Unscaled ray quantities: dx_u = dx*dc, (A'dy)_u = gray/dc
    double acc[NACC] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = blockIdx.x * RED_THREADS + threadIdx.x; j < n; j += gridDim.x * RED_THREADS) {
        const double dxs = xp[j] - x0[j];
        const double dxu = dxs * dc[j];
        const double lj = l[j], uj = u[j];
        acc[0] += c[j] * dxs;                                   // c'dx (scale-invariant)
        double vb = 0.0;                                        // ray must not push against finite bounds
        if (isfinite(lj) && dxu < 0.0) vb += dxu;
        if (isfinite(uj) && dxu > 0.0) vb += dxu;
        acc[1] += vb * vb;
        acc[3] += dxu * dxu;
        // dual ray: reduced cost of the ray r = -A'dy ; parts no finite bound can absorb are residual
        const double r = -gray[j];
        const double rpos = fmax(r, 0.0), rneg = fmin(r, 0.0);
        const double res = ((isfinite(lj) ? 0.0 : rpos) + (isfinite(uj) ? 0.0 : rneg)) / dc[j];
        acc[4] += res * res;
        acc[5] += (isfinite(lj) ? rpos * lj : 0.0) + (isfinite(uj) ? rneg * uj : 0.0);   // bound part of ray objective
    }
    block_reduce_store(acc, partials);
}
// rows: axray = A (xp - x0) (scaled; unscaled = /dr) ; dy = yp - y0 (unscaled = *dr)
__global__ void __launch_bounds__(RED_THREADS)
k_ray_rows(int m, const double* __restrict__ axray, const double* __restrict__ yp, const double* __restrict__ y0,
           const double* __restrict__ lc, const double* __restrict__ uc, const double* __restrict__ dr,
           double* __restrict__ partials) {
    double acc[NACC] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < m; i += gridDim.x * RED_THREADS) {
        const double a = axray[i] / dr[i];
        const double lo = lc[i], hi = uc[i];
        double v = 0.0;                                         // primal ray: A dx must stay inside the recession cone
        if (isfinite(lo) && a < 0.0) v += a;
        if (isfinite(hi) && a > 0.0) v += a;
        acc[0] += v * v;
        const double dys = yp[i] - y0[i];
        const double dyu = dys * dr[i];
        acc[1] += dyu * dyu;
        // dual ray objective (rows): dy+ * lc + dy- * uc (scale-invariant); infinite-side components are violations
        if (dys > 0.0) { if (isfinite(lo)) acc[2] += dys * lo; else acc[3] += dyu * dyu; }
        else if (dys < 0.0) { if (isfinite(hi)) acc[2] += dys * hi; else acc[3] += dyu * dyu; }
    }
    block_reduce_store(acc, partials);
}

// ---- small launch helpers used by build_csc / certificates --
__global__ void k_transpose_keys(const int* __restrict__ col, uint32_t nnz, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ perm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) { keys[i] = (uint64_t)(uint32_t)col[i]; perm[i] = i; }
}
__global__ void k_transpose_gather(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                                   const int* __restrict__ row_of, const double* __restrict__ val, uint32_t nnz,
                                   int* __restrict__ csc_idx, double* __restrict__ csc_val, int* __restrict__ cols_sorted,
                                   uint32_t* __restrict__ nnz_d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *nnz_d = nnz;
    if (i >= nnz) return;
    const uint32_t p = perm[i];
    csc_idx[i] = row_of[p];
    csc_val[i] = val[p];
    cols_sorted[i] = (int)keys[i];
}
__global__ void k_fill_ptr(const int* __restrict__ sorted_ids, const uint32_t* __restrict__ nnz_ptr, uint32_t nseg,
                           int* __restrict__ ptr, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nnz = *nnz_ptr;
    if (i > nnz || i > cap) return;
    const int prev = (i == 0) ? -1 : sorted_ids[i - 1];
    const int cur = (i == nnz) ? (int)nseg : sorted_ids[i];
    for (int r = prev + 1; r <= cur; ++r) ptr[r] = (int)i;
}
__global__ void k_diff(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}
void launch_transpose_keys(const int* col, uint32_t nnz, uint64_t* keys, uint32_t* perm, cudaStream_t st) {
    ELP_LAUNCH(k_transpose_keys, ceil_div(nnz, 256), 256, 0, st, col, nnz, keys, perm);
}
void launch_transpose_gather(const uint64_t* keys, const uint32_t* perm, const int* row_of, const double* val,
                             uint32_t nnz, int* csc_idx, double* csc_val, int* cols_sorted, uint32_t* nnz_d,
                             cudaStream_t st) {
    ELP_LAUNCH(k_transpose_gather, ceil_div(nnz, 256), 256, 0, st, keys, perm, row_of, val, nnz, csc_idx, csc_val,
               cols_sorted, nnz_d);
}
void launch_fill_ptr(const int* sorted_ids, const uint32_t* nnz_ptr, uint32_t nseg, int* ptr, uint32_t cap,
                     cudaStream_t st) {
    ELP_LAUNCH(k_fill_ptr, ceil_div((int64_t)cap + 1, 256), 256, 0, st, sorted_ids, nnz_ptr, nseg, ptr, cap);
}
void launch_diff(int n, const double* a, const double* b, double* out, cudaStream_t st) {
    if (n > 0) ELP_LAUNCH(k_diff, ceil_div(n, 256), 256, 0, st, n, a, b, out);
}

// ---- cross-GPU hand-off of the blocks written by peer stores ------------------------------------------------
// Every rank runs the same sequence of (signal, wait) pairs; the epoch lives in device memory so the pair can be
// replayed from a CUDA graph.  signal: runs after the producing kernel in stream order (its peer stores are complete),
// bumps this rank's epoch and publishes it in slot `rank` of every peer's flag row.  wait: spins (bounded) until all
// N slots of the local flag row have reached the local expectation.
struct PeerFlags {
    unsigned long long* row[8];   // row[r]: rank r's flag row (N slots), mapped into this process
};
__global__ void k_peer_signal(PeerFlags f, int nranks, int rank, unsigned long long* __restrict__ epoch) {
    if (threadIdx.x == 0) { *epoch += 1ull; __threadfence_system(); }
    __syncthreads();
    const unsigned long long e = *epoch;
    if ((int)threadIdx.x < nranks) {
        volatile unsigned long long* dst = f.row[threadIdx.x] + rank;
        *dst = e;
    }
    __threadfence_system();
}
__global__ void k_peer_wait(const unsigned long long* __restrict__ my_row, int nranks, unsigned long long* __restrict__ expect) {
    __shared__ unsigned long long want;
    if (threadIdx.x == 0) { *expect += 1ull; want = *expect; }
    __syncthreads();
    if ((int)threadIdx.x < nranks) {
        const volatile unsigned long long* src = my_row + threadIdx.x;
        unsigned long long spins = 0;
        while (*src < want) {
            if (++spins > (1ull << 28)) __trap();           // ~30 s: a peer died — fail instead of hanging the GPU
            __nanosleep(64);
        }
    }
    __threadfence_system();
}

// ---- which entries of the gathered vectors does each rank read? (sparse peer exchange) ------------------------
__global__ void k_mark_used(uint32_t nnz, const int* __restrict__ idx, unsigned char* __restrict__ used) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) used[idx[i]] = 1;                       // same value from every writer
}
// mask[j] = sum over ranks r of used_r[first + j] << r, own bit always set (the local kernels read the local copy)
__global__ void k_build_mask(int count, int first, size_t stride, int nranks, int rank, const unsigned char* __restrict__ used_all,
                             unsigned char* __restrict__ mask, unsigned long long* __restrict__ sent) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    unsigned mk = 1u << rank;
    for (int r = 0; r < nranks; ++r) mk |= (used_all[(size_t)r * stride + first + j] ? 1u : 0u) << r;
    mask[j] = (unsigned char)mk;
    atomicAdd(sent, (unsigned long long)(__popc(mk) - 1));
    if ((j & 3) == 0) {          // the same count per 32-byte line (4 entries): what the links actually carry
        unsigned line = mk;
        for (int q = 1; q < 4 && j + q < count; ++q)
            for (int r = 0; r < nranks; ++r) line |= (used_all[(size_t)r * stride + first + j + q] ? 1u : 0u) << r;
        atomicAdd(sent + 2, (unsigned long long)(__popc(line | (1u << rank)) - 1));
    }
}

// ---- kernels of the one-off exchange that builds each rank's column block of A ---------------------
__global__ void k_shift_idx(uint32_t nnz, int* __restrict__ idx, int offset) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) idx[i] += offset;
}
// lens[j] = length of local CSC column j0 + j (0 beyond the matrix)
__global__ void k_col_lens(const int* __restrict__ ptr, int j0, int count, int ncols, uint32_t* __restrict__ lens) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const int col = j0 + j;
    lens[j] = col < ncols ? (uint32_t)(ptr[col + 1] - ptr[col]) : 0u;
}
__global__ void k_sum_lens(int nranks, int nb, const uint32_t* __restrict__ lens, uint32_t* __restrict__ tot) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > nb) return;
    uint32_t s = 0;
    if (j < nb) for (int r = 0; r < nranks; ++r) s += lens[(size_t)r * nb + j];
    tot[j] = s;                                      // tot[nb] = 0: the scan leaves the total there
}
// column j of my block = concatenation over the source ranks (ascending = ascending global row) of their pieces
__global__ void k_merge_cols(int nranks, int nb, const uint32_t* __restrict__ lens, const uint32_t* __restrict__ srcptr,
                             const uint32_t* __restrict__ newptr, const uint32_t* __restrict__ roff,
                             const int* __restrict__ in_idx, const double* __restrict__ in_val,
                             int* __restrict__ out_ptr, int* __restrict__ out_idx, double* __restrict__ out_val) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > nb) return;
    out_ptr[j] = (int)newptr[j];
    if (j == nb) return;
    uint32_t dst = newptr[j];
    for (int r = 0; r < nranks; ++r) {
        const uint32_t len = lens[(size_t)r * nb + j];
        const uint32_t src = roff[r] + srcptr[(size_t)r * nb + j];
        for (uint32_t k = 0; k < len; ++k) { out_idx[dst + k] = in_idx[src + k]; out_val[dst + k] = in_val[src + k]; }
        dst += len;
    }
}

// ------------------------------------------------------------------------------------------------
// host-side solver object
//
// Partition (N ranks; N = 1 is the same code with the collectives compiled out of the picture):
//   rows     rank g owns the row block the caller hands it (balanced by non-zeros): CSR with GLOBAL column ids,
//            the slices of y / lc / uc.  y lives in a padded global layout y_full[N][mb] (mb = largest block).
//   columns  rank g owns the columns [g*nb, (g+1)*nb): the CSC of that column block over ALL rows (built once by
//            a grouped send/recv of the ranks' local CSC slices), the slices of x / c / l / u.
//   K1 (A'y + primal update) runs on the column block and gathers from y_full; it writes its x-bar block into
//   xbar_full, which is then all-gathered.  K2 (A x-bar + dual update) runs on the row block and gathers from
//   xbar_full; it updates its y block inside y_full, which is then all-gathered.  All work is sharded (nothing is
//   replicated), the exchange per iteration is 8 (n + m) bytes received per rank, and the scalar residuals of a
//   check travel in one 16-double allreduce.
// ------------------------------------------------------------------------------------------------
struct Pdlp {
    // single GPU: every persistent buffer of the handle is a piece of `arena` (one cudaMalloc / cudaFree instead of ~40),
    // the temporaries of the setup pieces of `scratch` (freed when the setup is over).  Declared first: destroyed last.
    Arena arena, scratch;
    int N = 1, rank = 0;
    int m = 0, mb = 0;          // local rows, padded rows per rank
    int n = 0, nb = 0, n0 = 0, nl = 0;   // global columns, columns per rank, my first column, my column count
    int64_t nnz = 0, nnzc = 0;  // entries of the row block / of the column block
    bool maximize = false;
    elp_options opt{};
    cudaStream_t st = nullptr;
    // matrix (scaled in place after setup)
    DevBuf<int> csr_ptr, csr_idx, csc_ptr, csc_idx;
    DevBuf<double> csr_val, csc_val;
    int Lr = 8, Lc = 4;          // lanes per row of the setup helper kernels
    SpmvPlan plan_r, plan_c;     // tile plans of the two iteration SpMVs (CSR rows / CSC columns)
    // vectors (scaled); column-side ones have nl entries, row-side ones m
    DevBuf<double> c, l, u, dc, x, x0, xp, gcol;
    DevBuf<double> lc, uc, dr, y0, yp, axbar, axp;
    DevBuf<double> xbar_full, y_full, xaux_full, yaux_full;   // [N*nb], [N*mb]: gathered vectors + scratch
    // peer-store exchange (N > 1, CUDA IPC): every rank's xbar_full / y_full / flag rows mapped here
    bool scatter = false;        // plain iterations: g = A'y accumulated by the dual kernel's scatter (single GPU only)
    SpmvPlan plan_s;             // tile plan of the scatter kernel (CSR rows)
    bool p2p = false;
    PeerOut x_out{}, y_out{};            // peers' xbar_full + n0, y_full + rank*mb
    PeerFlags xflags{}, yflags{};
    DevBuf<unsigned long long> flags;    // [2][8] flag rows (x, y) + [4] epochs/expectations
    DevBuf<unsigned char> xmask, ymask;  // sparse exchange: which ranks gather entry j of my x-bar / y block
    double exch_frac_x = 1.0, exch_frac_y = 1.0;   // fraction of the dense (N-1)-copy exchange that is actually sent
    std::vector<void*> ipc_opened;
    DevBuf<double> partials, scal;                              // scal: 2*NACC
    DevBuf<PdlpParams> params;
    // scalars
    double eta = 1.0, w = 1.0, w_init = 1.0, norm_b = 0.0, norm_c = 0.0, sigma_max = 0.0;
    int k = 0, total = 0, restarts = 0;
    double fpe0 = -1.0, fpe_prev = -1.0;
    double beta_artificial = 0.36;   // artificial restart once the epoch is this fraction of all iterations so far
    bool need_fpe0 = true;
    int status = ELP_STATUS_TIMEOUT;
    bool finished = false;
    double pobj = 0, dobj = 0, rel_pres = 0, rel_dres = 0, rel_gap = 0;
    int checks = 0;
    // graph of one chunk of (check_every-1) plain iterations
    cudaGraphExec_t graph = nullptr;
    int graph_len = 0;
    int64_t h2d = 0, d2h = 0;

    ~Pdlp() {
        if (graph) cudaGraphExecDestroy(graph);
        if (st) cudaStreamSynchronize(st);
        for (void* p : ipc_opened) cudaIpcCloseMemHandle(p);
        if (st) cudaStreamDestroy(st);
    }

    // Maps every rank's gathered buffers into this process (CUDA IPC) so that the iteration kernels can store their
    // block straight into all copies over NVLink.  Falls back to the NCCL all-gather when IPC is not available.
    void setup_peer_stores() {
        p2p = false;
        if (N <= 1 || N > 8 || env_int("ELP_PDLP_P2P", 1) == 0) return;
        flags.alloc(2 * 8 + 4);
        flags.zero(st);
        ELP_CUDA(cudaStreamSynchronize(st));
        struct Handles { cudaIpcMemHandle_t x, y, f; };
        static_assert(sizeof(Handles) % 8 == 0, "handle block must be a multiple of 8 bytes");
        std::vector<Handles> all(N);
        int ok = 1;
        if (cudaIpcGetMemHandle(&all[rank].x, xbar_full.p) != cudaSuccess) ok = 0;
        if (cudaIpcGetMemHandle(&all[rank].y, y_full.p) != cudaSuccess) ok = 0;
        if (cudaIpcGetMemHandle(&all[rank].f, flags.p) != cudaSuccess) ok = 0;
        cudaGetLastError();
        DevBuf<unsigned char> hbuf((size_t)N * sizeof(Handles));
        ELP_CUDA(cudaMemcpyAsync(hbuf.p + (size_t)rank * sizeof(Handles), &all[rank], sizeof(Handles), cudaMemcpyHostToDevice, st));
        comm_allgather_bytes(hbuf.p, sizeof(Handles), st);
        ELP_CUDA(cudaMemcpyAsync(all.data(), hbuf.p, (size_t)N * sizeof(Handles), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        std::vector<double*> xs(N), ys(N);
        std::vector<unsigned long long*> fs(N);
        for (int r = 0; r < N && ok; ++r) {
            if (r == rank) { xs[r] = xbar_full.p; ys[r] = y_full.p; fs[r] = flags.p; continue; }
            void *px = nullptr, *py = nullptr, *pf = nullptr;
            if (cudaIpcOpenMemHandle(&px, all[r].x, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            ipc_opened.push_back(px);
            if (cudaIpcOpenMemHandle(&py, all[r].y, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            ipc_opened.push_back(py);
            if (cudaIpcOpenMemHandle(&pf, all[r].f, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            ipc_opened.push_back(pf);
            xs[r] = (double*)px; ys[r] = (double*)py; fs[r] = (unsigned long long*)pf;
        }
        cudaGetLastError();
        // everyone or no one
        double okd = (double)ok;
        ELP_CUDA(cudaMemcpyAsync(scal.p, &okd, sizeof okd, cudaMemcpyHostToDevice, st));
        comm_allreduce_sum(scal.p, 1, st);
        ELP_CUDA(cudaMemcpyAsync(&okd, scal.p, sizeof okd, cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        if ((int)okd != N) {
            if (opt.verbose > 0 && rank == 0) fprintf(stderr, "[pdlp] CUDA IPC unavailable: NCCL all-gather exchange\n");
            return;
        }
        x_out.n = y_out.n = N;
        x_out.mask = y_out.mask = nullptr;
        for (int r = 0; r < N; ++r) {
            x_out.p[r] = xs[r] + n0;
            y_out.p[r] = ys[r] + (size_t)rank * mb;
            xflags.row[r] = fs[r];
            yflags.row[r] = fs[r] + 8;
        }
        p2p = true;
        if (env_int("ELP_PDLP_SPARSE_EXCHANGE", 1)) build_exchange_masks();
    }
    // Every rank marks the columns its row block references (K2 gathers x-bar there) and the rows its column block
    // references (K1 gathers y there); the marks are all-gathered and each rank keeps, for the entries it PRODUCES, the
    // set of ranks that read them.  The epilogues then store an entry only into those ranks' copies.
    void build_exchange_masks() {
        const size_t sx = (size_t)N * nb, sy = (size_t)N * mb;
        DevBuf<unsigned char> used_x((size_t)N * sx), used_y((size_t)N * sy);
        DevBuf<unsigned long long> sent(4);      // entries x-bar, entries y, 32-byte lines (both), unused
        used_x.zero(st); used_y.zero(st); sent.zero(st);
        if (nnz > 0) ELP_LAUNCH(k_mark_used, ceil_div(nnz, 256), 256, 0, st, (uint32_t)nnz, csr_idx.p, used_x.p + (size_t)rank * sx);
        if (nnzc > 0) ELP_LAUNCH(k_mark_used, ceil_div(nnzc, 256), 256, 0, st, (uint32_t)nnzc, csc_idx.p, used_y.p + (size_t)rank * sy);
        comm_allgather_bytes(used_x.p, sx, st);
        comm_allgather_bytes(used_y.p, sy, st);
        xmask.alloc((size_t)std::max(nl, 1)); ymask.alloc((size_t)std::max(m, 1));
        if (nl > 0) ELP_LAUNCH(k_build_mask, grid1(nl), 256, 0, st, nl, n0, sx, N, rank, used_x.p, xmask.p, sent.p);
        if (m > 0) ELP_LAUNCH(k_build_mask, grid1(m), 256, 0, st, m, rank * mb, sy, N, rank, used_y.p, ymask.p, sent.p + 1);
        // (k_build_mask adds its line count to sent[+2]: the y launch starts at sent + 1, so lines of y land in sent[3])
        unsigned long long h[4] = {0, 0, 0, 0};
        ELP_CUDA(cudaMemcpyAsync(h, sent.p, sizeof h, cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        exch_frac_x = nl > 0 ? (double)h[0] / ((double)nl * (N - 1)) : 1.0;
        exch_frac_y = m > 0 ? (double)h[1] / ((double)m * (N - 1)) : 1.0;
        // all ranks take the same decision.  The masks cost a byte load and a predicate per store, and the links move
        // partial 32-byte lines as dearly as full ones, so the decision is taken on LINES kept: measured at N = 2, 60 %
        // kept (C5) -> 4 % slower, 92-100 % (C4) -> 3 % slower; at N = 8, 23 % kept in long runs (C5) -> 1.9x faster.
        // A random matrix keeps 1 - exp(-nnz_local / n) of the entries (C4 at N = 8: 46 %) but 91 % of the lines.
        double tot[4] = {(double)h[2] + (double)h[3], ((double)((nl + 3) / 4) + (double)((m + 3) / 4)) * (N - 1), 0, 0};
        ELP_CUDA(cudaMemcpyAsync(scal.p, tot, 2 * sizeof(double), cudaMemcpyHostToDevice, st));
        comm_allreduce_sum(scal.p, 2, st);
        ELP_CUDA(cudaMemcpyAsync(tot, scal.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        const double kept = tot[1] > 0 ? tot[0] / tot[1] : 1.0;
        const bool use = env_int("ELP_PDLP_SPARSE_EXCHANGE", 1) == 2 || kept < 0.50;
        if (use) {
            x_out.mask = xmask.p;
            y_out.mask = ymask.p;
        }
        if (opt.verbose > 0 || env_int("ELP_PDLP_DEBUG", 0))
            fprintf(stderr, "[pdlp] rank %d sparse exchange: x-bar %.1f %%, y %.1f %% of the dense peer stores; 32-byte lines, all ranks %.1f %% -> %s\n",
                    rank, 100.0 * exch_frac_x, 100.0 * exch_frac_y, 100.0 * kept, use ? "masked stores" : "dense stores");
    }
    void barrier_stream() {              // all ranks have finished everything they queued before this point
        if (N > 1) comm_allreduce_sum(scal.p + NACC, 1, st);
    }

    int grid1(int count) const { return std::max(1, ceil_div(count, 256)); }
    double* xbar() { return xbar_full.p + n0; }              // my block of the gathered x-bar
    double* y() { return y_full.p + (size_t)rank * mb; }     // my block of the gathered y
    void gather_x(double* full) { if (N > 1) comm_allgather(full, (size_t)nb, st); }
    void gather_y(double* full) { if (N > 1) comm_allgather(full, (size_t)mb, st); }

    void reduce_to(double* out_dev) {
        ELP_LAUNCH(k_final_reduce, 1, RED_THREADS, 0, st, partials.p, RED_BLOCKS, out_dev);
    }
    // sum over all ranks of a k_* reduction that was just launched into `partials`; returns NACC values
    void fetch_scalars(double* host) {
        reduce_to(scal.p);
        if (N > 1) comm_allreduce_sum(scal.p, NACC, st);
        ELP_CUDA(cudaMemcpyAsync(host, scal.p, NACC * sizeof(double), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        d2h += NACC * sizeof(double);
    }

    // out[m] = A_rows * v_full      (my rows; v_full in the flat column layout)
    void spmv_rows(const double* v_full, double* out) {
        launch_spmv(plan_r, m, csr_ptr.p, csr_idx.p, csr_val.p, v_full, store_epi(out), st);
    }
    // out[nl] = (A' * v_full)_mycols  (v_full in the padded row layout)
    void spmv_cols(const double* v_full, double* out) {
        launch_spmv(plan_c, nl, csc_ptr.p, csc_idx.p, csc_val.p, v_full, store_epi(out), st);
    }

    template <int MODE>
    void rowstat(int L, int nrows, const int* ptr, const int* idx, const double* val, const double* s_self,
                 const double* s_other, double* out) {
        if (nrows <= 0) return;
        const int grid = ceil_div((int64_t)nrows * L, SPMV_THREADS);
        switch (L) {
            case 1:  ELP_LAUNCH((rowstat_kernel<1, MODE>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other, out); break;
            case 2:  ELP_LAUNCH((rowstat_kernel<2, MODE>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other, out); break;
            case 4:  ELP_LAUNCH((rowstat_kernel<4, MODE>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other, out); break;
            case 8:  ELP_LAUNCH((rowstat_kernel<8, MODE>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other, out); break;
            case 16: ELP_LAUNCH((rowstat_kernel<16, MODE>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other, out); break;
            default: ELP_LAUNCH((rowstat_kernel<32, MODE>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other, out); break;
        }
    }
    void scale_vals(int L, int nrows, const int* ptr, const int* idx, double* val, const double* s_self,
                    const double* s_other) {
        if (nrows <= 0) return;
        const int grid = ceil_div((int64_t)nrows * L, SPMV_THREADS);
        switch (L) {
            case 1:  ELP_LAUNCH((scale_vals_kernel<1>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other); break;
            case 2:  ELP_LAUNCH((scale_vals_kernel<2>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other); break;
            case 4:  ELP_LAUNCH((scale_vals_kernel<4>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other); break;
            case 8:  ELP_LAUNCH((scale_vals_kernel<8>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other); break;
            case 16: ELP_LAUNCH((scale_vals_kernel<16>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other); break;
            default: ELP_LAUNCH((scale_vals_kernel<32>), grid, SPMV_THREADS, 0, st, nrows, ptr, idx, val, s_self, s_other); break;
        }
    }
    void expand_rows(int L, int nrows, const int* ptr, int* row_of) {
        if (nrows <= 0) return;
        const int grid = ceil_div((int64_t)nrows * L, SPMV_THREADS);
        switch (L) {
            case 1:  ELP_LAUNCH((expand_rows_kernel<1>), grid, SPMV_THREADS, 0, st, nrows, ptr, row_of); break;
            case 2:  ELP_LAUNCH((expand_rows_kernel<2>), grid, SPMV_THREADS, 0, st, nrows, ptr, row_of); break;
            case 4:  ELP_LAUNCH((expand_rows_kernel<4>), grid, SPMV_THREADS, 0, st, nrows, ptr, row_of); break;
            case 8:  ELP_LAUNCH((expand_rows_kernel<8>), grid, SPMV_THREADS, 0, st, nrows, ptr, row_of); break;
            case 16: ELP_LAUNCH((expand_rows_kernel<16>), grid, SPMV_THREADS, 0, st, nrows, ptr, row_of); break;
            default: ELP_LAUNCH((expand_rows_kernel<32>), grid, SPMV_THREADS, 0, st, nrows, ptr, row_of); break;
        }
    }

    // ---- setup ---------------------------------------------------------------------------------
    void setup(int m_, int n_, const int32_t* row_ptr, const int32_t* col_idx, const double* vals,
               const int8_t* sense, const double* rhs, const double* c_h, int maximize_, const double* lb,
               const double* ub, const elp_options& o, bool dist_, int64_t nnz_device = -1) {
        // nnz_device >= 0: row_ptr / col_idx / vals are DEVICE arrays (a model assembled by elp_model_assemble)
        m = m_; n = n_; maximize = maximize_ != 0; opt = o;
        WallTimer dbg_t;
        const bool dbg = getenv("ELP_PDLP_DEBUG") != nullptr;
        auto mark = [&](const char* what) {
            if (!dbg) return;
            cudaStreamSynchronize(st);
            fprintf(stderr, "[pdlp setup] %-16s %9.3f ms\n", what, dbg_t.ms());
        };
        const bool dist = dist_ && comm().active;
        N = dist ? comm().nranks : 1;
        rank = dist ? comm().rank : 0;
        ELP_REQUIRE(m >= 0 && n > 0, "pdlp: bad shape %d x %d", m, n);
        nnz = nnz_device >= 0 ? nnz_device : (m > 0 ? row_ptr[m] : 0);
        ELP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (N == 1 && env_int("ELP_PDLP_ARENA", 1)) {   // (peer-mapped buffers of the distributed path need their own allocations)
            arena.reserve((size_t)24 * nnz + (size_t)4 * (m + n) + (size_t)8 * ((size_t)10 * n + (size_t)9 * m) + (2u << 20));
            scratch.reserve((size_t)52 * nnz + (size_t)16 * ((size_t)n + m) + (8u << 20));
        }
        ArenaScope persistent(arena.base ? &arena : nullptr);
        partials.alloc((size_t)RED_BLOCKS * NACC); scal.alloc(2 * NACC); params.alloc(1);
        partials.zero(st);
        // block sizes: multiples of 4 keep every block 32-byte aligned (16-byte-granular bulk copies)
        nb = (ceil_div(n, N) + 3) & ~3;
        n0 = std::min(n, rank * nb);
        nl = std::max(0, std::min(n, (rank + 1) * nb) - n0);
        {
            double mx = (double)m;
            if (N > 1) {
                ELP_CUDA(cudaMemcpyAsync(scal.p, &mx, sizeof mx, cudaMemcpyHostToDevice, st));
                comm_allreduce_max(scal.p, 1, st);
                ELP_CUDA(cudaMemcpyAsync(&mx, scal.p, sizeof mx, cudaMemcpyDeviceToHost, st));
                ELP_CUDA(cudaStreamSynchronize(st));
            }
            mb = (std::max(1, (int)mx) + 3) & ~3;
        }
        ELP_REQUIRE((int64_t)N * mb < 0x7fffffffll && (int64_t)N * nb < 0x7fffffffll, "pdlp: problem too large for int32 ids");

        csr_ptr.alloc(m + 1 + SPMV_PTR_PAD); csr_idx.alloc(nnz + SPMV_PAD); csr_val.alloc(nnz + SPMV_PAD);
        csr_idx.zero(st); csr_val.zero(st); csr_ptr.zero(st);
        // epilogue operands are staged by 16-byte-granular bulk copies: SPMV_VPAD doubles of slack each
        const size_t np = (size_t)std::max(nl, 1) + SPMV_VPAD, mp = (size_t)std::max(m, 1) + SPMV_VPAD;
        c.alloc(np); l.alloc(np); u.alloc(np); dc.alloc(np); x.alloc(np); x0.alloc(np); xp.alloc(np); gcol.alloc(np);
        lc.alloc(mp); uc.alloc(mp); dr.alloc(mp); y0.alloc(mp); yp.alloc(mp); axbar.alloc(mp); axp.alloc(mp);
        xbar_full.alloc((size_t)N * nb + SPMV_VPAD); xaux_full.alloc((size_t)N * nb + SPMV_VPAD);
        y_full.alloc((size_t)N * mb + SPMV_VPAD); yaux_full.alloc((size_t)N * mb + SPMV_VPAD);
        for (DevBuf<double>* b : {&c, &l, &u, &dc, &x, &x0, &xp, &gcol, &lc, &uc, &dr, &y0, &yp, &axbar, &axp, &xbar_full,
                                  &xaux_full, &y_full, &yaux_full})
            b->zero(st);

        if (nnz_device >= 0) {
            ELP_CUDA(cudaStreamSynchronize(0));           // the assembly ran on the default stream
            ELP_CUDA(cudaMemcpyAsync(csr_ptr.p, row_ptr, ((size_t)m + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st));
            if (nnz) {
                ELP_CUDA(cudaMemcpyAsync(csr_idx.p, col_idx, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToDevice, st));
                ELP_CUDA(cudaMemcpyAsync(csr_val.p, vals, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice, st));
            }
        } else {
            if (m == 0) { int z = 0; csr_ptr.upload(&z, 1, st); }
            else csr_ptr.upload(row_ptr, m + 1, st);
            csr_idx.upload(col_idx, nnz, st);
            csr_val.upload(vals, nnz, st);
            h2d += (int64_t)(m + 1) * 4 + nnz * 12;
        }
        c.upload(c_h + n0, nl, st); l.upload(lb + n0, nl, st); u.upload(ub + n0, nl, st);
        h2d += (int64_t)nl * 24;
        {
            ArenaScope tmp(scratch.base ? &scratch : g_arena);
            DevBuf<int8_t> sense_d(std::max(m, 1));
            DevBuf<double> rhs_d(std::max(m, 1));
            sense_d.upload(sense, m, st); rhs_d.upload(rhs, m, st);
            h2d += (int64_t)m * 9;
            if (m > 0) ELP_LAUNCH(k_row_bounds, grid1(m), 256, 0, st, m, sense_d.p, rhs_d.p, lc.p, uc.p);
            ELP_CUDA(cudaStreamSynchronize(st));
        }
        if (maximize && nl > 0) ELP_LAUNCH(k_scale_scalar, grid1(nl), 256, 0, st, nl, c.p, -1.0);

        mark("upload");
        build_column_block();
        mark("column block");
        Lr = pick_helper_lanes(nnz, m);
        Lc = pick_helper_lanes(nnzc, nl);
        plan_r = plan_spmv(nnz, m, 4);
        plan_c = plan_spmv(nnzc, nl, 5);
        {
            int mode = opt.transpose;
            if (const int e = env_int("ELP_PDLP_TRANSPOSE", 0)) mode = e;      // sweeps / A-B runs
            ELP_REQUIRE(mode >= ELP_TRANSPOSE_AUTO && mode <= ELP_TRANSPOSE_SCATTER, "pdlp: bad transpose mode %d", mode);
            ELP_REQUIRE(!(mode == ELP_TRANSPOSE_SCATTER && N > 1), "pdlp: the scatter formulation is single-GPU only");
            scatter = N == 1 && mode == ELP_TRANSPOSE_SCATTER && m > 0;
            if (const char* e = getenv("ELP_PDLP_BETA_ART")) beta_artificial = atof(e);       // experiments
            plan_s = plan_spmv(nnz, m, 4, 0, env_int("ELP_SPMV_SCAT_STAGES", 2));
        }

        // unscaled norms for the relative termination test
        double h[NACC];
        ELP_LAUNCH(k_bound_norm, RED_BLOCKS, RED_THREADS, 0, st, m, lc.p, uc.p, partials.p);
        fetch_scalars(h);
        norm_b = std::sqrt(h[0]);
        ELP_LAUNCH(k_dot2, RED_BLOCKS, RED_THREADS, 0, st, nl, c.p, c.p, partials.p);
        fetch_scalars(h);
        norm_c = std::sqrt(h[0]);

        mark("norms");
        scale_problem();
        mark("scaling");
        estimate_sigma_max();
        mark("power iteration");
        eta = sigma_max > 0 ? 0.998 / sigma_max : 1.0;
        // initial primal weight from the scaled data: ||c|| / ||b||
        ELP_LAUNCH(k_bound_norm, RED_BLOCKS, RED_THREADS, 0, st, m, lc.p, uc.p, partials.p);
        fetch_scalars(h);
        const double nbn = std::sqrt(h[0]);
        ELP_LAUNCH(k_dot2, RED_BLOCKS, RED_THREADS, 0, st, nl, c.p, c.p, partials.p);
        fetch_scalars(h);
        const double ncn = std::sqrt(h[0]);
        w_init = (nbn > 1e-10 && ncn > 1e-10) ? ncn / nbn : 1.0;
        setup_peer_stores();
        reset();
        mark("reset");
        ELP_CUDA(cudaStreamSynchronize(st));
        scratch.release();                           // the setup's temporaries, in one cudaFree
    }

    // CSC of my column block over all rows, row ids in the padded y_full layout.  Every rank transposes its own row
    // block with the stable radix sort that backs the assembly, then the column slices travel to their owners in
    // one grouped send/recv and are merged source by source (= ascending global row).
    void build_column_block() {
        ArenaScope tmp(scratch.base ? &scratch : g_arena);       // temporaries; the CSC arrays below go to the handle's arena
        // ---- local transpose: CSC of my row block over all n columns ----------------------------------
        DevBuf<int> lptr((size_t)n + 1), lidx(std::max<int64_t>(nnz, 1));
        DevBuf<double> lval(std::max<int64_t>(nnz, 1));
        if (nnz == 0) {
            lptr.zero(st);
        } else {
            DevBuf<uint64_t> keys(nnz);
            DevBuf<uint32_t> perm(nnz), nnz_d(1);
            DevBuf<int> row_of(nnz), cols_sorted(nnz);
            RadixSortWorkspace ws;
            expand_rows(pick_helper_lanes(nnz, m), m, csr_ptr.p, row_of.p);
            launch_transpose_keys(csr_idx.p, (uint32_t)nnz, keys.p, perm.p, st);
            radix_sort_pairs(keys.p, perm.p, nnz, bit_length_u64((uint64_t)n - 1), ws, st);
            launch_transpose_gather(keys.p, perm.p, row_of.p, csr_val.p, (uint32_t)nnz, lidx.p, lval.p, cols_sorted.p,
                                    nnz_d.p, st);
            launch_fill_ptr(cols_sorted.p, nnz_d.p, (uint32_t)n, lptr.p, (uint32_t)nnz, st);
            if (rank > 0) ELP_LAUNCH(k_shift_idx, ceil_div(nnz, 256), 256, 0, st, (uint32_t)nnz, lidx.p, rank * mb);
            ELP_CUDA(cudaStreamSynchronize(st));
        }
        if (N == 1) {
            nnzc = nnz;
            ArenaScope keep(arena.base ? &arena : nullptr);
            csc_ptr.alloc((size_t)nl + 1 + SPMV_PTR_PAD); csc_idx.alloc(nnzc + SPMV_PAD); csc_val.alloc(nnzc + SPMV_PAD);
            csc_ptr.zero(st); csc_idx.zero(st); csc_val.zero(st);
            ELP_CUDA(cudaMemcpyAsync(csc_ptr.p, lptr.p, ((size_t)n + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st));
            if (nnz) {
                ELP_CUDA(cudaMemcpyAsync(csc_idx.p, lidx.p, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToDevice, st));
                ELP_CUDA(cudaMemcpyAsync(csc_val.p, lval.p, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice, st));
            }
            ELP_CUDA(cudaStreamSynchronize(st));
            return;
        }
        // ---- how much goes where ----------------------------------------------------------------------
        std::vector<int> cut(N + 1);
        for (int g = 0; g <= N; ++g) {
            const int col = std::min(n, g * nb);
            ELP_CUDA(cudaMemcpyAsync(&cut[g], lptr.p + col, sizeof(int), cudaMemcpyDeviceToHost, st));
        }
        ELP_CUDA(cudaStreamSynchronize(st));
        DevBuf<double> cnt_d((size_t)N * N);
        std::vector<double> cnt((size_t)N * N, 0.0);
        for (int g = 0; g < N; ++g) cnt[(size_t)rank * N + g] = (double)(cut[g + 1] - cut[g]);
        ELP_CUDA(cudaMemcpyAsync(cnt_d.p + (size_t)rank * N, cnt.data() + (size_t)rank * N, N * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
        comm_allgather(cnt_d.p, (size_t)N, st);
        ELP_CUDA(cudaMemcpyAsync(cnt.data(), cnt_d.p, (size_t)N * N * sizeof(double), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        std::vector<uint32_t> roff(N + 1, 0);
        for (int r = 0; r < N; ++r) roff[r + 1] = roff[r] + (uint32_t)cnt[(size_t)r * N + rank];   // what r sends me
        const int64_t total_in = roff[N];
        // ---- lens of every column slice I send, then the grouped exchange -------------------------------
        DevBuf<uint32_t> lens_out((size_t)N * nb), lens_in((size_t)N * nb + 1), srcptr((size_t)N * nb), newptr((size_t)nb + 1);
        DevBuf<uint32_t> roff_d(N + 1);
        DevBuf<int> in_idx(std::max<int64_t>(total_in, 1));
        DevBuf<double> in_val(std::max<int64_t>(total_in, 1));
        for (int g = 0; g < N; ++g)
            ELP_LAUNCH(k_col_lens, ceil_div(nb, 256), 256, 0, st, lptr.p, g * nb, nb, n, lens_out.p + (size_t)g * nb);
        comm_group_start();
        for (int g = 0; g < N; ++g) {
            const size_t cntg = (size_t)(cut[g + 1] - cut[g]);
            comm_send(lens_out.p + (size_t)g * nb, (size_t)nb * sizeof(uint32_t), g, st);
            comm_send(lidx.p + cut[g], cntg * sizeof(int), g, st);
            comm_send(lval.p + cut[g], cntg * sizeof(double), g, st);
            const size_t cin = (size_t)(roff[g + 1] - roff[g]);
            comm_recv(lens_in.p + (size_t)g * nb, (size_t)nb * sizeof(uint32_t), g, st);
            comm_recv(in_idx.p + roff[g], cin * sizeof(int), g, st);
            comm_recv(in_val.p + roff[g], cin * sizeof(double), g, st);
        }
        comm_group_end();
        // ---- merge ---------------------------------------------------------------------------------------
        ScanWorkspace sw;
        ELP_LAUNCH(k_sum_lens, ceil_div(nb + 1, 256), 256, 0, st, N, nb, lens_in.p, newptr.p);
        exclusive_scan_u32(newptr.p, (size_t)nb + 1, sw, st);
        ELP_CUDA(cudaMemcpyAsync(srcptr.p, lens_in.p, (size_t)N * nb * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        for (int r = 0; r < N; ++r) exclusive_scan_u32(srcptr.p + (size_t)r * nb, (size_t)nb, sw, st);
        roff_d.upload(roff.data(), N + 1, st);
        nnzc = total_in;
        csc_ptr.alloc((size_t)nb + 1 + SPMV_PTR_PAD); csc_idx.alloc(nnzc + SPMV_PAD); csc_val.alloc(nnzc + SPMV_PAD);
        csc_ptr.zero(st); csc_idx.zero(st); csc_val.zero(st);
        ELP_LAUNCH(k_merge_cols, ceil_div(nb + 1, 256), 256, 0, st, N, nb, lens_in.p, srcptr.p, newptr.p, roff_d.p,
                   in_idx.p, in_val.p, csc_ptr.p, csc_idx.p, csc_val.p);
        ELP_CUDA(cudaStreamSynchronize(st));
    }

    void scale_problem() {
        ArenaScope tmp(scratch.base ? &scratch : g_arena);
        const int ruiz = opt.ruiz_iters >= 0 ? opt.ruiz_iters : 10;
        // dr_full = yaux_full (padded row layout), dc_full = xaux_full (flat column layout) during scaling
        double* dr_full = yaux_full.p;
        double* dc_full = xaux_full.p;
        double* drl = dr_full + (size_t)rank * mb;
        double* dcl = dc_full + n0;
        ELP_LAUNCH(k_fill, grid1(N * mb), 256, 0, st, N * mb, dr_full, 1.0);
        ELP_LAUNCH(k_fill, grid1(N * nb), 256, 0, st, N * nb, dc_full, 1.0);
        DevBuf<double> rstat(std::max(m, 1)), cstat(std::max(nl, 1));
        for (int it = 0; it < ruiz + 1; ++it) {
            const bool pc = it == ruiz;   // last pass: Pock-Chambolle (alpha = 1) with L1 norms
            if (!pc) {
                rowstat<0>(Lr, m, csr_ptr.p, csr_idx.p, csr_val.p, drl, dc_full, rstat.p);
                rowstat<0>(Lc, nl, csc_ptr.p, csc_idx.p, csc_val.p, dcl, dr_full, cstat.p);
            } else {
                rowstat<1>(Lr, m, csr_ptr.p, csr_idx.p, csr_val.p, drl, dc_full, rstat.p);
                rowstat<1>(Lc, nl, csc_ptr.p, csc_idx.p, csc_val.p, dcl, dr_full, cstat.p);
            }
            if (m > 0) ELP_LAUNCH(k_ruiz_update, grid1(m), 256, 0, st, m, drl, rstat.p);
            if (nl > 0) ELP_LAUNCH(k_ruiz_update, grid1(nl), 256, 0, st, nl, dcl, cstat.p);
            gather_y(dr_full);
            gather_x(dc_full);
        }
        scale_vals(Lr, m, csr_ptr.p, csr_idx.p, csr_val.p, drl, dc_full);
        scale_vals(Lc, nl, csc_ptr.p, csc_idx.p, csc_val.p, dcl, dr_full);
        if (m > 0) ELP_CUDA(cudaMemcpyAsync(dr.p, drl, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (nl > 0) ELP_CUDA(cudaMemcpyAsync(dc.p, dcl, (size_t)nl * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (nl > 0) {
            ELP_LAUNCH(k_mul, grid1(nl), 256, 0, st, nl, c.p, dc.p);
            ELP_LAUNCH(k_div, grid1(nl), 256, 0, st, nl, l.p, dc.p);
            ELP_LAUNCH(k_div, grid1(nl), 256, 0, st, nl, u.p, dc.p);
        }
        if (m > 0) {
            ELP_LAUNCH(k_mul, grid1(m), 256, 0, st, m, lc.p, dr.p);
            ELP_LAUNCH(k_mul, grid1(m), 256, 0, st, m, uc.p, dr.p);
        }
    }

    void estimate_sigma_max() {
        sigma_max = 0.0;
        double h[NACC];
        // has the matrix any entry at all?
        double tot = (double)nnz;
        if (N > 1) {
            ELP_CUDA(cudaMemcpyAsync(scal.p, &tot, sizeof tot, cudaMemcpyHostToDevice, st));
            comm_allreduce_sum(scal.p, 1, st);
            ELP_CUDA(cudaMemcpyAsync(&tot, scal.p, sizeof tot, cudaMemcpyDeviceToHost, st));
            ELP_CUDA(cudaStreamSynchronize(st));
        }
        if (tot == 0.0) return;
        // power iteration on A'A: v in xaux_full (my block at n0), A v in yaux_full (my block), A'(A v) in gcol
        double* v = xaux_full.p + n0;
        double* av = yaux_full.p + (size_t)rank * mb;
        ELP_CUDA(cudaMemsetAsync(xaux_full.p, 0, (size_t)N * nb * sizeof(double), st));
        ELP_CUDA(cudaMemsetAsync(yaux_full.p, 0, (size_t)N * mb * sizeof(double), st));
        if (nl > 0) ELP_LAUNCH(k_pseudo_random, grid1(nl), 256, 0, st, nl, v, 12345u + 7919u * (uint32_t)n0);
        ELP_LAUNCH(k_dot2, RED_BLOCKS, RED_THREADS, 0, st, nl, v, v, partials.p);
        fetch_scalars(h);
        if (nl > 0) ELP_LAUNCH(k_scale_scalar, grid1(nl), 256, 0, st, nl, v, 1.0 / std::sqrt(h[0]));
        double s = 1.0;
        for (int it = 0; it < 60; ++it) {
            gather_x(xaux_full.p);
            spmv_rows(xaux_full.p, av);
            gather_y(yaux_full.p);
            spmv_cols(yaux_full.p, gcol.p);
            ELP_LAUNCH(k_dot2, RED_BLOCKS, RED_THREADS, 0, st, nl, gcol.p, gcol.p, partials.p);
            fetch_scalars(h);
            const double nrm = std::sqrt(h[0]);
            if (!(nrm > 0.0)) { s = 0.0; break; }
            const double s_new = std::sqrt(nrm);
            if (nl > 0) {
                ELP_CUDA(cudaMemcpyAsync(v, gcol.p, (size_t)nl * sizeof(double), cudaMemcpyDeviceToDevice, st));
                ELP_LAUNCH(k_scale_scalar, grid1(nl), 256, 0, st, nl, v, 1.0 / nrm);
            }
            const bool conv = std::fabs(s_new - s) <= 1e-4 * s_new;
            s = s_new;
            if (conv && it >= 10) break;
        }
        sigma_max = s;
    }

    void reset() {
        if (nl > 0) ELP_LAUNCH(k_init_x, grid1(nl), 256, 0, st, nl, x.p, l.p, u.p);
        ELP_CUDA(cudaMemcpyAsync(x0.p, x.p, (size_t)std::max(nl, 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
        ELP_CUDA(cudaMemcpyAsync(xp.p, x.p, (size_t)std::max(nl, 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
        y_full.zero(st); y0.zero(st); yp.zero(st); axbar.zero(st); axp.zero(st); xbar_full.zero(st);
        gcol.zero(st);                             // = A'y for y = 0 (scatter formulation)
        w = w_init; k = 0; total = 0; restarts = 0; fpe0 = -1; fpe_prev = -1; need_fpe0 = true;
        status = ELP_STATUS_TIMEOUT; finished = false; checks = 0;
        push_params();
        barrier_stream();          // nobody stores into a peer's copy before that peer has cleared it
        ELP_CUDA(cudaStreamSynchronize(st));
    }

    void push_params() {
        PdlpParams p{eta / w, eta * w, k, 0};
        // pageable source is copied to a staging buffer before the call returns, so a stack object is fine
        ELP_CUDA(cudaMemcpyAsync(params.p, &p, sizeof p, cudaMemcpyHostToDevice, st));
    }

    // ---- iteration pieces ------------------------------------------------------------------------
    template <bool CHECK>
    void primal_step(int it, bool exchange = true) {
        if (!CHECK && scatter) {            // g = A'y is already in gcol (left there by the dual kernel's scatter)
            const int grid = std::max(1, std::min(ceil_div(nl / 2, 256), kNumSMs * 8));
            ELP_LAUNCH(k_primal_from_g, grid, 256, 0, st, nl, gcol.p, c.p, l.p, u.p, x0.p, x.p, xbar(), params.p, it);
            return;
        }
        const bool direct = !CHECK && exchange && p2p;
        PrimalEpi<CHECK> epi{nullptr, c.p, l.p, u.p, x0.p, x.p, xbar(), xp.p, params.p, it, direct ? x_out : PeerOut{{}, 0, nullptr}};
        launch_spmv(plan_c, nl, csc_ptr.p, csc_idx.p, csc_val.p, y_full.p, epi, st);
        if (!exchange) return;
        if (direct) {
            ELP_LAUNCH(k_peer_signal, 1, 32, 0, st, xflags, N, rank, flags.p + 16);
            ELP_LAUNCH(k_peer_wait, 1, 32, 0, st, flags.p, N, flags.p + 17);
        } else {
            gather_x(xbar_full.p);
        }
    }
    template <bool CHECK>
    void dual_step(int it, bool exchange = true) {
        if (!CHECK && scatter) {            // A x-bar + dual update, then g += val * y_new over the row's entries
            DualEpi<false, true> epi{gcol.p, lc.p, uc.p, y0.p, y(), yp.p, axbar.p, params.p, it, PeerOut{{}, 0, nullptr}};
            launch_spmv(plan_s, m, csr_ptr.p, csr_idx.p, csr_val.p, xbar_full.p, epi, st);
            return;
        }
        const bool direct = !CHECK && exchange && p2p;
        DualEpi<CHECK> epi{nullptr, lc.p, uc.p, y0.p, y(), yp.p, axbar.p, params.p, it, direct ? y_out : PeerOut{{}, 0, nullptr}};
        launch_spmv(plan_r, m, csr_ptr.p, csr_idx.p, csr_val.p, xbar_full.p, epi, st);
        if (!exchange || CHECK) return;                      // a check iteration does not change y here
        if (direct) {
            ELP_LAUNCH(k_peer_signal, 1, 32, 0, st, yflags, N, rank, flags.p + 18);
            ELP_LAUNCH(k_peer_wait, 1, 32, 0, st, flags.p + 8, N, flags.p + 19);
        } else {
            gather_y(y_full.p);
        }
    }
    int kernels_per_iter() const { return (m > 0 ? 1 : 0) + (nl > 0 ? 1 : 0) + (p2p ? 4 : 0); }

    void plain_iterations(int count) {
        if (count <= 0) return;
        const bool want_graph = opt.use_graph != 0 && count == opt.check_every - 1 && count > 1;
        if (want_graph) {
            if (!graph || graph_len != count) {
                if (graph) { cudaGraphExecDestroy(graph); graph = nullptr; }
                cudaGraph_t g = nullptr;
                const int64_t before = g_launches.load();
                ELP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                try {
                    for (int i = 0; i < count; ++i) { primal_step<false>(i); dual_step<false>(i); }
                } catch (...) {
                    cudaStreamEndCapture(st, &g);
                    if (g) cudaGraphDestroy(g);
                    throw;
                }
                ELP_CUDA(cudaStreamEndCapture(st, &g));
                g_launches.store(before);   // capture does not execute; launches are counted per replay
                ELP_CUDA(cudaGraphInstantiate(&graph, g, 0));
                cudaGraphDestroy(g);
                graph_len = count;
            }
            ELP_CUDA(cudaGraphLaunch(graph, st));
            g_launches.fetch_add((int64_t)count * kernels_per_iter());
        } else {
            for (int i = 0; i < count; ++i) { primal_step<false>(i); dual_step<false>(i); }
        }
        k += count;
        total += count;
        push_params();
    }

    // One check iteration: computes T(z) with side products, evaluates KKT + fixed-point error, then
    // either restarts from T(z) or completes the Halpern step.  Returns true when the solve is over.
    bool check_iteration() {
        double h[2 * NACC];
        primal_step<true>(0);                      // xbar (gathered), xp
        dual_step<true>(0);                        // yp, axbar; y unchanged
        // A xp on my rows (needs everyone's xp) and A' yp on my columns (needs everyone's yp)
        if (nl > 0) ELP_CUDA(cudaMemcpyAsync(xaux_full.p + n0, xp.p, (size_t)nl * sizeof(double), cudaMemcpyDeviceToDevice, st));
        gather_x(xaux_full.p);
        spmv_rows(xaux_full.p, axp.p);
        if (m > 0) ELP_CUDA(cudaMemcpyAsync(yaux_full.p + (size_t)rank * mb, yp.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, st));
        gather_y(yaux_full.p);
        spmv_cols(yaux_full.p, gcol.p);
        ELP_LAUNCH(k_check_rows, RED_BLOCKS, RED_THREADS, 0, st, m, axp.p, axbar.p, y(), yp.p, y0.p, lc.p, uc.p, dr.p,
                   partials.p);
        reduce_to(scal.p);
        ELP_LAUNCH(k_check_cols, RED_BLOCKS, RED_THREADS, 0, st, nl, gcol.p, c.p, l.p, u.p, x.p, xp.p, x0.p, dc.p,
                   partials.p);
        reduce_to(scal.p + NACC);
        if (N > 1) comm_allreduce_sum(scal.p, 2 * NACC, st);       // the scalar residuals: ONE collective
        ELP_CUDA(cudaMemcpyAsync(h, scal.p, 2 * NACC * sizeof(double), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        d2h += 2 * NACC * sizeof(double);
        ++checks;
        const double* hr = h;
        const double* hc = h + NACC;

        const double tau = eta / w, sigma = eta * w;
        const double fpe2 = hc[1] / tau + 2.0 * hr[1] + hr[2] / sigma;
        const double fpe = std::sqrt(std::max(fpe2, 0.0));
        pobj = hc[3];
        dobj = hr[4] + hc[4];
        rel_pres = std::sqrt(hr[0]) / (1.0 + norm_b);
        rel_dres = std::sqrt(hc[0]) / (1.0 + norm_c);
        rel_gap = std::fabs(pobj - dobj) / (1.0 + std::fabs(pobj) + std::fabs(dobj));
        const double eps = opt.eps_rel;
        if (opt.verbose > 0 && rank == 0)
            fprintf(stderr, "[pdlp] it %7d k %6d pres %.3e dres %.3e gap %.3e pobj %.10e fpe %.3e w %.3e restarts %d\n",
                    total + 1, k, rel_pres, rel_dres, rel_gap, maximize ? -pobj : pobj, fpe, w, restarts);
        ++total;   // this check iteration is a full PDHG iteration
        if (!std::isfinite(fpe) || !std::isfinite(pobj)) {
            status = ELP_STATUS_NUMFAILURE; finished = true; return true;
        }
        // The gap is tested at eps/4: |p - d| <= (eps/4)(1 + |p| + |d|) keeps the objective within eps RELATIVE of
        // the optimum (north_star: "objective within 1e-6 relative"); PDLP's plain gap test allows ~2 eps.
        if (rel_pres <= eps && rel_dres <= eps && rel_gap <= 0.25 * eps) {
            status = ELP_STATUS_OPTIMAL; finished = true; return true;
        }
        if (k > 0 && (checks % 4 == 0) && detect_infeasible()) { finished = true; return true; }

        if (need_fpe0) { fpe0 = fpe; need_fpe0 = false; }
        bool restart = false;
        if (k > 0) {
            if (fpe <= 0.2 * fpe0) restart = true;
            else if (fpe <= 0.8 * fpe0 && fpe_prev >= 0 && fpe > fpe_prev) restart = true;
            else if ((double)k >= beta_artificial * (double)total) restart = true;
        }
        fpe_prev = fpe;
        const int span = std::max(std::max(nl, m), 1);
        if (restart) {
            const double ddx = std::sqrt(hc[2]), ddy = std::sqrt(hr[3]);
            if (ddx > 1e-10 && ddy > 1e-10) w = std::exp(0.5 * std::log(ddy / ddx) + 0.5 * std::log(w));
            ELP_LAUNCH(k_restart, grid1(span), 256, 0, st, nl, x.p, x0.p, xp.p, m, y(), y0.p, yp.p);
            k = 0; ++restarts; need_fpe0 = true; fpe_prev = -1;
        } else {
            const double wk = (k + 1.0) / (k + 2.0);
            ELP_LAUNCH(k_halpern_finish, grid1(span), 256, 0, st, nl, x.p, xbar(), x0.p, m, y(), yp.p, y0.p, wk);
            ++k;
        }
        gather_y(y_full.p);                        // y changed: everyone needs the new blocks
        if (scatter) spmv_cols(y_full.p, gcol.p);  // ... and the scatter formulation needs g = A'y of the new y
        push_params();
        return false;
    }

    // Farkas-type certificates from the displacement (xp - x0, yp - y0); see oracle/pdlp_ref.py::_certificate
    bool detect_infeasible() {
        double h[2 * NACC];
        const double tol = 1e-6;
        // ray vectors in the scratch gathered buffers; their images in axp / gcol (both free at this point)
        double* dxl = xaux_full.p + n0;
        double* dyl = yaux_full.p + (size_t)rank * mb;
        launch_diff(nl, xp.p, x0.p, dxl, st);
        launch_diff(m, yp.p, y0.p, dyl, st);
        gather_x(xaux_full.p);
        gather_y(yaux_full.p);
        spmv_rows(xaux_full.p, axp.p);
        spmv_cols(yaux_full.p, gcol.p);
        ELP_LAUNCH(k_ray_rows, RED_BLOCKS, RED_THREADS, 0, st, m, axp.p, yp.p, y0.p, lc.p, uc.p, dr.p, partials.p);
        reduce_to(scal.p);
        ELP_LAUNCH(k_ray_cols, RED_BLOCKS, RED_THREADS, 0, st, nl, xp.p, x0.p, c.p, l.p, u.p, gcol.p, dc.p, partials.p);
        reduce_to(scal.p + NACC);
        if (N > 1) comm_allreduce_sum(scal.p, 2 * NACC, st);
        ELP_CUDA(cudaMemcpyAsync(h, scal.p, 2 * NACC * sizeof(double), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        d2h += 2 * NACC * sizeof(double);
        const double* hr = h;
        const double* hc = h + NACC;
        // primal infeasibility: dual ray with positive objective and vanishing residual
        const double ray_obj = hr[2] + hc[5];
        const double ny = std::sqrt(hr[1]);
        if (ny > 1e-12 && ray_obj > 0.0) {
            const double res = std::sqrt(hc[4]) + std::sqrt(hr[3]);
            if (res / ray_obj <= tol) { status = ELP_STATUS_INFEASIBLE; return true; }
        }
        // dual infeasibility (primal unbounded): primal ray with negative cost staying feasible
        const double cdx = hc[0];
        const double nx = std::sqrt(hc[3]);
        if (nx > 1e-12 && cdx < 0.0) {
            const double viol = std::sqrt(hr[0]) + std::sqrt(hc[1]);
            if (viol / (-cdx) <= tol) { status = ELP_STATUS_UNBOUNDED; return true; }
        }
        return false;
    }

    void run(int max_new_iters, elp_stats* stats) {
        WallTimer wall;
        const int64_t launches0 = g_launches.load();
        cudaEvent_t e0, e1;
        ELP_CUDA(cudaEventCreate(&e0)); ELP_CUDA(cudaEventCreate(&e1));
        ELP_CUDA(cudaEventRecord(e0, st));
        const int ce = std::max(2, opt.check_every);
        opt.check_every = ce;
        const int limit_total = opt.max_iter > 0 ? opt.max_iter : 2000000;
        int budget = max_new_iters > 0 ? max_new_iters : limit_total;
        if (finished && status != ELP_STATUS_TIMEOUT) budget = 0;
        finished = false;
        while (budget > 0 && total < limit_total) {
            // epoch-relative schedule: a check whenever k % check_every == 0 (k = 0 right after a restart)
            if (k % ce == 0) {
                if (check_iteration()) break;
                --budget;
                if (N == 1 && opt.time_limit_s > 0 && wall.ms() > opt.time_limit_s * 1e3) break;   // ranks must agree: no wall-clock exit when distributed
                continue;
            }
            int cnt = ce - (k % ce);
            cnt = std::min(cnt, std::min(budget, limit_total - total));
            plain_iterations(cnt);
            budget -= cnt;
        }
        ELP_CUDA(cudaEventRecord(e1, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        ELP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (stats) {
            stats->status = status;
            stats->method_used = ELP_METHOD_PDLP;
            stats->iterations = total;
            stats->restarts = restarts;
            stats->primal_obj = maximize ? -pobj : pobj;
            stats->dual_obj = maximize ? -dobj : dobj;
            stats->rel_primal_res = rel_pres;
            stats->rel_dual_res = rel_dres;
            stats->rel_gap = rel_gap;
            stats->solve_ms = ms;
            stats->total_ms = wall.ms();
            stats->kernel_launches = g_launches.load() - launches0;
            stats->h2d_bytes = h2d;
            stats->d2h_bytes = d2h;
        }
    }

    // x: all n columns (gathered from the ranks' blocks); y: my rows
    void solution(double* x_h, double* y_h, double* obj) {
        if (x_h) {
            if (nl > 0) ELP_LAUNCH(k_unscale, grid1(nl), 256, 0, st, nl, xp.p, dc.p, 1.0, xaux_full.p + n0);
            gather_x(xaux_full.p);
            xaux_full.download(x_h, n, st);
            ELP_CUDA(cudaStreamSynchronize(st));
            d2h += (int64_t)n * 8;
        }
        if (y_h && m > 0) {
            ELP_LAUNCH(k_unscale, grid1(m), 256, 0, st, m, yp.p, dr.p, maximize ? -1.0 : 1.0, axp.p);
            axp.download(y_h, m, st);
            ELP_CUDA(cudaStreamSynchronize(st));
            d2h += (int64_t)m * 8;
        }
        if (obj) *obj = maximize ? -pobj : pobj;
    }

    // Times the two iteration kernels as they run in a solve: `reps` iterations (primal kernel, dual kernel, no
    // collective) from a mid-solve iterate, one CUDA event between any two launches, so that the two figures add up to
    // the iteration time.  Leaves the solver reset.
    void probe_step(int reps, double* ms_primal, double* ms_dual) {
        reps = std::max(1, std::min(reps, 512));
        reset();
        run(3 * std::max(2, opt.check_every), nullptr);
        std::vector<cudaEvent_t> ev(2 * (size_t)reps + 1);
        for (auto& e : ev) ELP_CUDA(cudaEventCreate(&e));
        for (int i = 0; i < 3; ++i) { primal_step<false>(i, false); dual_step<false>(i, false); }
        for (int i = 0; i < reps; ++i) {
            ELP_CUDA(cudaEventRecord(ev[2 * i], st));
            primal_step<false>(i, false);
            ELP_CUDA(cudaEventRecord(ev[2 * i + 1], st));
            dual_step<false>(i, false);
        }
        ELP_CUDA(cudaEventRecord(ev[2 * reps], st));
        ELP_CUDA(cudaStreamSynchronize(st));
        double tp = 0.0, td = 0.0;
        for (int i = 0; i < reps; ++i) {
            float ms = 0;
            ELP_CUDA(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
            tp += ms;
            ELP_CUDA(cudaEventElapsedTime(&ms, ev[2 * i + 1], ev[2 * i + 2]));
            td += ms;
        }
        for (auto& e : ev) cudaEventDestroy(e);
        if (ms_primal) *ms_primal = tp / reps;
        if (ms_dual) *ms_dual = td / reps;
        reset();
    }

    void probe_spmv(int reps, double* ms_csr, double* ms_csc) {
        cudaEvent_t e0, e1;
        ELP_CUDA(cudaEventCreate(&e0)); ELP_CUDA(cudaEventCreate(&e1));
        float ms = 0;
        for (int i = 0; i < 3; ++i) { spmv_rows(xbar_full.p, axp.p); }
        ELP_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) spmv_rows(xbar_full.p, axp.p);
        ELP_CUDA(cudaEventRecord(e1, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        ELP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_csr) *ms_csr = ms / reps;
        for (int i = 0; i < 3; ++i) spmv_cols(y_full.p, gcol.p);
        ELP_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) spmv_cols(y_full.p, gcol.p);
        ELP_CUDA(cudaEventRecord(e1, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        ELP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_csc) *ms_csc = ms / reps;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
};

// ---- entry points used by abi.cu -----------------------------------------------------------------
Pdlp* pdlp_create(int m, int n, const int32_t* row_ptr, const int32_t* col_idx, const double* vals,
                  const int8_t* sense, const double* rhs, const double* c, int maximize, const double* lb,
                  const double* ub, const elp_options& opt, bool dist, elp_stats* stats, int64_t nnz_device) {
    WallTimer t;
    auto* p = new Pdlp();
    try {
        p->setup(m, n, row_ptr, col_idx, vals, sense, rhs, c, maximize, lb, ub, opt, dist, nnz_device);
    } catch (...) {
        delete p;
        throw;
    }
    if (stats) { stats->setup_ms = t.ms(); stats->h2d_bytes = p->h2d; }
    return p;
}
void pdlp_run(Pdlp* p, int max_new_iters, elp_stats* stats) { p->run(max_new_iters, stats); }
void pdlp_reset(Pdlp* p) { p->reset(); }
void pdlp_solution(Pdlp* p, double* x, double* y, double* obj) { p->solution(x, y, obj); }
void pdlp_probe(Pdlp* p, int reps, double* a, double* b) { p->probe_spmv(reps, a, b); }
void pdlp_probe_step(Pdlp* p, int reps, double* a, double* b) { p->probe_step(reps, a, b); }
int pdlp_transpose(Pdlp* p) { return p->scatter ? ELP_TRANSPOSE_SCATTER : ELP_TRANSPOSE_GATHER; }
void pdlp_destroy(Pdlp* p) { delete p; }

// plain SpMV + feasibility re-check (S4: /root/reference/R/class.R:533-540, R/utils.R:167-171)
__global__ void k_compare_tol(int m, const double* __restrict__ lhs, const double* __restrict__ rhs,
                              const int8_t* __restrict__ sense, double tol, uint8_t* __restrict__ ok) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double a = lhs[i], b = rhs[i];
    bool f;
    switch (sense[i]) {
        case 0: f = (a + tol <= b) || (a - tol <= b); break;           // "<="  any(lhs +/- tol <= rhs)
        case 1: f = (a + tol >= b) || (a - tol >= b); break;           // ">="
        case 2: f = fabs(a - b) <= tol; break;                          // "=="
        case 3: f = (a + tol < b) || (a - tol < b); break;             // "<"  stays strict (match.fun(dir))
        default: f = (a + tol > b) || (a - tol > b); break;            // ">"
    }
    ok[i] = f ? 1 : 0;
}

void spmv_host(int m, int n, const int32_t* row_ptr, const int32_t* col_idx, const double* vals, const double* x,
               double* out, const int8_t* sense, const double* rhs, double tol, uint8_t* feasible) {
    cudaStream_t st = 0;
    const int64_t nnz = m > 0 ? row_ptr[m] : 0;
    if (m == 0) return;
    Arena arena;                 // eight buffers, one allocation
    arena.reserve((size_t)12 * (nnz + SPMV_PAD) + (size_t)8 * std::max(n, 1) + (size_t)30 * (m + 8) + 16 * 512);
    ArenaScope scope(arena.base ? &arena : nullptr);
    DevBuf<int> ptr(m + 1 + SPMV_PTR_PAD), idx(nnz + SPMV_PAD);
    DevBuf<double> val(nnz + SPMV_PAD), xd(std::max(n, 1)), od(m);
    idx.zero(st); val.zero(st); ptr.zero(st);
    ptr.upload(row_ptr, m + 1, st); idx.upload(col_idx, nnz, st); val.upload(vals, nnz, st); xd.upload(x, n, st);
    // one thread per row: the row sum is formed in index order, bit-identical to a scalar loop
    launch_spmv(plan_spmv(nnz, m, 0, 1), m, ptr.p, idx.p, val.p, xd.p, store_epi(od.p), st);
    if (out) od.download(out, m, st);
    if (feasible) {
        DevBuf<int8_t> sd(m);
        DevBuf<double> rd(m);
        DevBuf<uint8_t> fd(m);
        sd.upload(sense, m, st); rd.upload(rhs, m, st);
        ELP_LAUNCH(k_compare_tol, ceil_div(m, 256), 256, 0, st, m, od.p, rd.p, sd.p, tol, fd.p);
        fd.download(feasible, m, st);
        ELP_CUDA(cudaStreamSynchronize(st));
        return;
    }
    ELP_CUDA(cudaStreamSynchronize(st));
}

}  // namespace elp
