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
// Multi-GPU (row partition): each rank owns a row block (its CSR, the CSC of the same
// block, its slice of y) and a replica of x.  K1 splits into [partial A_g'y_g] -> ncclAllReduce ->
// [primal update]; scalar partial sums ride in the tail of the allreduce buffer at check points.
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

template <class Epi>
__global__ void __launch_bounds__(256) apply_epi_kernel(int n, const double* __restrict__ g, Epi epi) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < n) epi.apply(j, g[j], epi.preload_global(j));
}

// ---- epilogues of the SpMV (spmv.cuh): in(i) names the operand vectors the producer stages next to the
// matrix stream, preload() picks this row's operands out of the stage, apply() runs after the row sum ------
struct StoreEpi {
    static constexpr int NIN = 0;
    double* out;
    struct Pre {};
    __device__ __forceinline__ const double* in(int) const { return nullptr; }
    __device__ __forceinline__ Pre preload(const double*, int, int) const { return Pre{}; }
    __device__ __forceinline__ Pre preload_global(int) const { return Pre{}; }
    __device__ __forceinline__ void apply(int r, double s, const Pre&) const { out[r] = s; }
};

// primal half of T(z) + reflection + Halpern combine.  g = (A'y)_j
template <bool CHECK>
struct PrimalEpi {
    static constexpr int NIN = CHECK ? 4 : 5;
    const double* __restrict__ c;
    const double* __restrict__ l;
    const double* __restrict__ u;
    const double* __restrict__ x0;
    double* __restrict__ x;
    double* __restrict__ xbar;
    double* __restrict__ xp;
    const PdlpParams* __restrict__ P;
    int it;
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
    __device__ __forceinline__ void apply(int j, double g, const Pre& p) const {
        const double tau = P->tau;
        const double xpj = fmin(fmax(p.x - tau * (p.c - g), p.l), p.u);
        const double xb = 2.0 * xpj - p.x;
        xbar[j] = xb;
        if (CHECK) {
            xp[j] = xpj;
        } else {
            const double k = (double)(P->k_base + it);
            const double w = (k + 1.0) / (k + 2.0);
            x[j] = w * xb + (1.0 - w) * p.x0;
        }
    }
};

// dual half.  ax = (A xbar)_i
template <bool CHECK>
struct DualEpi {
    static constexpr int NIN = CHECK ? 3 : 4;
    const double* __restrict__ lc;
    const double* __restrict__ uc;
    const double* __restrict__ y0;
    double* __restrict__ y;
    double* __restrict__ yp;
    double* __restrict__ axbar;
    const PdlpParams* __restrict__ P;
    int it;
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
    __device__ __forceinline__ void apply(int i, double ax, const Pre& p) const {
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
            y[i] = w * (2.0 * ypi - p.y) + (1.0 - w) * p.y0;
        }
    }
};

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

// ------------------------------------------------------------------------------------------------
// host-side solver object
// ------------------------------------------------------------------------------------------------
struct Pdlp {
    int m = 0, n = 0;           // local rows, global columns
    int64_t nnz = 0;
    bool dist = false;
    bool maximize = false;
    elp_options opt{};
    cudaStream_t st = nullptr;
    // matrix (scaled in place after setup)
    DevBuf<int> csr_ptr, csr_idx, csc_ptr, csc_idx;
    DevBuf<double> csr_val, csc_val;
    int Lr = 8, Lc = 4;          // lanes per row of the setup helper kernels
    SpmvPlan plan_r, plan_c;     // tile plans of the two iteration SpMVs (CSR rows / CSC columns)
    // vectors (scaled)
    DevBuf<double> c, l, u, lc, uc, dr, dc;
    DevBuf<double> x, x0, xbar, xp, y, y0, yp, axbar, axp, gbuf;   // gbuf: n + NACC (allreduce buffer)
    DevBuf<double> ray_dx, ray_dy, ray_ax, ray_g;                    // certificate scratch (lazy)
    DevBuf<double> partials, scal;                                   // scal: 2*NACC
    DevBuf<PdlpParams> params;
    // scalars
    double eta = 1.0, w = 1.0, w_init = 1.0, norm_b = 0.0, norm_c = 0.0, sigma_max = 0.0;
    int k = 0, total = 0, restarts = 0;
    double fpe0 = -1.0, fpe_prev = -1.0;
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
        if (st) cudaStreamDestroy(st);
    }

    int grid1(int count) const { return std::max(1, ceil_div(count, 256)); }

    void reduce_to(double* out_dev) {
        ELP_LAUNCH(k_final_reduce, 1, RED_THREADS, 0, st, partials.p, RED_BLOCKS, out_dev);
    }
    // sum over (all ranks of) a k_* reduction that was just launched into `partials`; returns NACC values
    void fetch_scalars(double* host, bool row_partitioned) {
        reduce_to(scal.p);
        if (dist && row_partitioned) comm_allreduce_sum(scal.p, NACC, st);
        ELP_CUDA(cudaMemcpyAsync(host, scal.p, NACC * sizeof(double), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        d2h += NACC * sizeof(double);
    }

    // y_out[m] = A * v      (local rows)
    void spmv_rows(const double* v, double* out) {
        launch_spmv(plan_r, m, csr_ptr.p, csr_idx.p, csr_val.p, v, StoreEpi{out}, st);
    }
    // out[n] = A' * v  (summed over ranks)
    void spmv_cols(const double* v, double* out) {
        launch_spmv(plan_c, n, csc_ptr.p, csc_idx.p, csc_val.p, v, StoreEpi{out}, st);
        if (dist) comm_allreduce_sum(out, n, st);
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

    // ---- setup ---------------------------------------------------------------------------------
    void setup(int m_, int n_, const int32_t* row_ptr, const int32_t* col_idx, const double* vals,
               const int8_t* sense, const double* rhs, const double* c_h, int maximize_, const double* lb,
               const double* ub, const elp_options& o, bool dist_) {
        m = m_; n = n_; dist = dist_ && comm().active; maximize = maximize_ != 0; opt = o;
        ELP_REQUIRE(m >= 0 && n > 0, "pdlp: bad shape %d x %d", m, n);
        nnz = m > 0 ? row_ptr[m] : 0;
        ELP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        csr_ptr.alloc(m + 1 + SPMV_PTR_PAD); csr_idx.alloc(nnz + SPMV_PAD); csr_val.alloc(nnz + SPMV_PAD);
        csc_ptr.alloc(n + 1 + SPMV_PTR_PAD); csc_idx.alloc(nnz + SPMV_PAD); csc_val.alloc(nnz + SPMV_PAD);
        csr_idx.zero(st); csr_val.zero(st); csc_idx.zero(st); csc_val.zero(st); csr_ptr.zero(st); csc_ptr.zero(st);
        // epilogue operands are staged by 16-byte-granular bulk copies: SPMV_VPAD doubles of slack each
        const size_t np = (size_t)n + SPMV_VPAD, mp = (size_t)std::max(m, 1) + SPMV_VPAD;
        c.alloc(np); l.alloc(np); u.alloc(np); dc.alloc(np);
        lc.alloc(mp); uc.alloc(mp); dr.alloc(mp);
        x.alloc(np); x0.alloc(np); xbar.alloc(np); xp.alloc(np); gbuf.alloc(n + NACC);
        y.alloc(mp); y0.alloc(mp); yp.alloc(mp);
        axbar.alloc(mp); axp.alloc(mp);
        c.zero(st); l.zero(st); u.zero(st); lc.zero(st); uc.zero(st); x.zero(st); x0.zero(st);
        partials.alloc((size_t)RED_BLOCKS * NACC); scal.alloc(2 * NACC); params.alloc(1);
        partials.zero(st);

        if (m == 0) { int z = 0; csr_ptr.upload(&z, 1, st); }
        else csr_ptr.upload(row_ptr, m + 1, st);
        csr_idx.upload(col_idx, nnz, st);
        csr_val.upload(vals, nnz, st);
        c.upload(c_h, n, st); l.upload(lb, n, st); u.upload(ub, n, st);
        h2d += (int64_t)(m + 1) * 4 + nnz * 12 + (int64_t)n * 24;
        {
            DevBuf<int8_t> sense_d(std::max(m, 1));
            DevBuf<double> rhs_d(std::max(m, 1));
            sense_d.upload(sense, m, st); rhs_d.upload(rhs, m, st);
            h2d += (int64_t)m * 9;
            if (m > 0) ELP_LAUNCH(k_row_bounds, grid1(m), 256, 0, st, m, sense_d.p, rhs_d.p, lc.p, uc.p);
            ELP_CUDA(cudaStreamSynchronize(st));
        }
        if (maximize) ELP_LAUNCH(k_scale_scalar, grid1(n), 256, 0, st, n, c.p, -1.0);
        Lr = pick_helper_lanes(nnz, m);
        Lc = pick_helper_lanes(nnz, n);
        plan_r = plan_spmv(nnz, m, 4);
        plan_c = plan_spmv(nnz, n, 5);

        build_csc();
        // unscaled norms for the relative termination test
        double h[NACC];
        ELP_LAUNCH(k_bound_norm, RED_BLOCKS, RED_THREADS, 0, st, m, lc.p, uc.p, partials.p);
        fetch_scalars(h, true);
        norm_b = std::sqrt(h[0]);
        ELP_LAUNCH(k_dot2, RED_BLOCKS, RED_THREADS, 0, st, n, c.p, c.p, partials.p);
        fetch_scalars(h, false);
        norm_c = std::sqrt(h[0]);

        scale_problem();
        estimate_sigma_max();
        eta = sigma_max > 0 ? 0.998 / sigma_max : 1.0;
        // initial primal weight from the scaled data: ||c|| / ||b||
        ELP_LAUNCH(k_bound_norm, RED_BLOCKS, RED_THREADS, 0, st, m, lc.p, uc.p, partials.p);
        fetch_scalars(h, true);
        const double nb = std::sqrt(h[0]);
        ELP_LAUNCH(k_dot2, RED_BLOCKS, RED_THREADS, 0, st, n, c.p, c.p, partials.p);
        fetch_scalars(h, false);
        const double nc = std::sqrt(h[0]);
        w_init = (nb > 1e-10 && nc > 1e-10) ? nc / nb : 1.0;
        reset();
    }

    void build_csc() {
        // transpose of the local block with the same stable sort that backs the assembly
        if (nnz == 0) { csc_ptr.zero(st); return; }
        DevBuf<uint64_t> keys(nnz);
        DevBuf<uint32_t> perm(nnz), nnz_d(1);
        DevBuf<int> row_of(nnz), cols_sorted(nnz);
        RadixSortWorkspace ws;
        {
            const int grid = ceil_div((int64_t)m * Lr, SPMV_THREADS);
            switch (Lr) {
                case 1:  ELP_LAUNCH((expand_rows_kernel<1>), grid, SPMV_THREADS, 0, st, m, csr_ptr.p, row_of.p); break;
                case 2:  ELP_LAUNCH((expand_rows_kernel<2>), grid, SPMV_THREADS, 0, st, m, csr_ptr.p, row_of.p); break;
                case 4:  ELP_LAUNCH((expand_rows_kernel<4>), grid, SPMV_THREADS, 0, st, m, csr_ptr.p, row_of.p); break;
                case 8:  ELP_LAUNCH((expand_rows_kernel<8>), grid, SPMV_THREADS, 0, st, m, csr_ptr.p, row_of.p); break;
                case 16: ELP_LAUNCH((expand_rows_kernel<16>), grid, SPMV_THREADS, 0, st, m, csr_ptr.p, row_of.p); break;
                default: ELP_LAUNCH((expand_rows_kernel<32>), grid, SPMV_THREADS, 0, st, m, csr_ptr.p, row_of.p); break;
            }
        }
        launch_transpose_keys(csr_idx.p, (uint32_t)nnz, keys.p, perm.p, st);
        radix_sort_pairs(keys.p, perm.p, nnz, bit_length_u64((uint64_t)n - 1), ws, st);
        launch_transpose_gather(keys.p, perm.p, row_of.p, csr_val.p, (uint32_t)nnz, csc_idx.p, csc_val.p, cols_sorted.p,
                                nnz_d.p, st);
        launch_fill_ptr(cols_sorted.p, nnz_d.p, (uint32_t)n, csc_ptr.p, (uint32_t)nnz, st);
        ELP_CUDA(cudaStreamSynchronize(st));
    }

    void scale_problem() {
        const int ruiz = opt.ruiz_iters >= 0 ? opt.ruiz_iters : 10;
        if (m > 0) ELP_LAUNCH(k_fill, grid1(m), 256, 0, st, m, dr.p, 1.0);
        ELP_LAUNCH(k_fill, grid1(n), 256, 0, st, n, dc.p, 1.0);
        if (nnz == 0) return;
        DevBuf<double> rstat(std::max(m, 1)), cstat(n);
        for (int it = 0; it < ruiz + 1; ++it) {
            const bool pc = it == ruiz;   // last pass: Pock-Chambolle (alpha = 1) with L1 norms
            if (!pc) {
                rowstat<0>(Lr, m, csr_ptr.p, csr_idx.p, csr_val.p, dr.p, dc.p, rstat.p);
                rowstat<0>(Lc, n, csc_ptr.p, csc_idx.p, csc_val.p, dc.p, dr.p, cstat.p);
                if (dist) comm_allreduce_max(cstat.p, n, st);
            } else {
                rowstat<1>(Lr, m, csr_ptr.p, csr_idx.p, csr_val.p, dr.p, dc.p, rstat.p);
                rowstat<1>(Lc, n, csc_ptr.p, csc_idx.p, csc_val.p, dc.p, dr.p, cstat.p);
                if (dist) comm_allreduce_sum(cstat.p, n, st);
            }
            ELP_LAUNCH(k_ruiz_update, grid1(m), 256, 0, st, m, dr.p, rstat.p);
            ELP_LAUNCH(k_ruiz_update, grid1(n), 256, 0, st, n, dc.p, cstat.p);
        }
        scale_vals(Lr, m, csr_ptr.p, csr_idx.p, csr_val.p, dr.p, dc.p);
        scale_vals(Lc, n, csc_ptr.p, csc_idx.p, csc_val.p, dc.p, dr.p);
        ELP_LAUNCH(k_mul, grid1(n), 256, 0, st, n, c.p, dc.p);
        ELP_LAUNCH(k_div, grid1(n), 256, 0, st, n, l.p, dc.p);
        ELP_LAUNCH(k_div, grid1(n), 256, 0, st, n, u.p, dc.p);
        if (m > 0) {
            ELP_LAUNCH(k_mul, grid1(m), 256, 0, st, m, lc.p, dr.p);
            ELP_LAUNCH(k_mul, grid1(m), 256, 0, st, m, uc.p, dr.p);
        }
    }

    void estimate_sigma_max() {
        sigma_max = 0.0;
        if (nnz == 0 || m == 0) return;
        double h[NACC];
        // power iteration on A'A, vector in xbar, A v in axbar, A'(A v) in gbuf
        ELP_LAUNCH(k_pseudo_random, grid1(n), 256, 0, st, n, xbar.p, 12345u);
        ELP_LAUNCH(k_dot2, RED_BLOCKS, RED_THREADS, 0, st, n, xbar.p, xbar.p, partials.p);
        fetch_scalars(h, false);
        ELP_LAUNCH(k_scale_scalar, grid1(n), 256, 0, st, n, xbar.p, 1.0 / std::sqrt(h[0]));
        double s = 1.0;
        for (int it = 0; it < 60; ++it) {
            spmv_rows(xbar.p, axbar.p);
            spmv_cols(axbar.p, gbuf.p);
            ELP_LAUNCH(k_dot2, RED_BLOCKS, RED_THREADS, 0, st, n, gbuf.p, gbuf.p, partials.p);
            fetch_scalars(h, false);
            const double nrm = std::sqrt(h[0]);
            if (!(nrm > 0.0)) { s = 0.0; break; }
            const double s_new = std::sqrt(nrm);
            ELP_CUDA(cudaMemcpyAsync(xbar.p, gbuf.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
            ELP_LAUNCH(k_scale_scalar, grid1(n), 256, 0, st, n, xbar.p, 1.0 / nrm);
            const bool conv = std::fabs(s_new - s) <= 1e-4 * s_new;
            s = s_new;
            if (conv && it >= 10) break;
        }
        sigma_max = s;
    }

    void reset() {
        ELP_LAUNCH(k_init_x, grid1(n), 256, 0, st, n, x.p, l.p, u.p);
        ELP_CUDA(cudaMemcpyAsync(x0.p, x.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        y.zero(st); y0.zero(st); yp.zero(st); axbar.zero(st); axp.zero(st);
        ELP_CUDA(cudaMemcpyAsync(xp.p, x.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        w = w_init; k = 0; total = 0; restarts = 0; fpe0 = -1; fpe_prev = -1; need_fpe0 = true;
        status = ELP_STATUS_TIMEOUT; finished = false; checks = 0;
        push_params();
        ELP_CUDA(cudaStreamSynchronize(st));
    }

    void push_params() {
        PdlpParams p{eta / w, eta * w, k, 0};
        // pageable source is copied to a staging buffer before the call returns, so a stack object is fine
        ELP_CUDA(cudaMemcpyAsync(params.p, &p, sizeof p, cudaMemcpyHostToDevice, st));
    }

    // ---- iteration pieces ------------------------------------------------------------------------
    template <bool CHECK>
    void primal_step(int it) {
        PrimalEpi<CHECK> epi{c.p, l.p, u.p, x0.p, x.p, xbar.p, xp.p, params.p, it};
        if (!dist) {
            launch_spmv(plan_c, n, csc_ptr.p, csc_idx.p, csc_val.p, y.p, epi, st);
        } else {
            launch_spmv(plan_c, n, csc_ptr.p, csc_idx.p, csc_val.p, y.p, StoreEpi{gbuf.p}, st);
            comm_allreduce_sum(gbuf.p, n, st);
            ELP_LAUNCH((apply_epi_kernel<PrimalEpi<CHECK>>), grid1(n), 256, 0, st, n, gbuf.p, epi);
        }
    }
    template <bool CHECK>
    void dual_step(int it) {
        DualEpi<CHECK> epi{lc.p, uc.p, y0.p, y.p, yp.p, axbar.p, params.p, it};
        launch_spmv(plan_r, m, csr_ptr.p, csr_idx.p, csr_val.p, xbar.p, epi, st);
    }
    int kernels_per_iter() const { return (m > 0 ? 1 : 0) + (dist ? 2 : 1); }

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
        double hr[NACC], hc[NACC];
        primal_step<true>(0);
        dual_step<true>(0);
        spmv_rows(xp.p, axp.p);
        ELP_LAUNCH(k_check_rows, RED_BLOCKS, RED_THREADS, 0, st, m, axp.p, axbar.p, y.p, yp.p, y0.p, lc.p, uc.p, dr.p,
                   partials.p);
        // row sums travel in the tail of the A'y buffer so a distributed check costs one allreduce
        reduce_to(gbuf.p + n);
        launch_spmv(plan_c, n, csc_ptr.p, csc_idx.p, csc_val.p, yp.p, StoreEpi{gbuf.p}, st);
        if (dist) comm_allreduce_sum(gbuf.p, (size_t)n + NACC, st);
        ELP_LAUNCH(k_check_cols, RED_BLOCKS, RED_THREADS, 0, st, n, gbuf.p, c.p, l.p, u.p, x.p, xp.p, x0.p, dc.p,
                   partials.p);
        reduce_to(scal.p);
        ELP_CUDA(cudaMemcpyAsync(hr, gbuf.p + n, NACC * sizeof(double), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaMemcpyAsync(hc, scal.p, NACC * sizeof(double), cudaMemcpyDeviceToHost, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        d2h += 2 * NACC * sizeof(double);
        ++checks;

        const double tau = eta / w, sigma = eta * w;
        const double fpe2 = hc[1] / tau + 2.0 * hr[1] + hr[2] / sigma;
        const double fpe = std::sqrt(std::max(fpe2, 0.0));
        pobj = hc[3];
        dobj = hr[4] + hc[4];
        rel_pres = std::sqrt(hr[0]) / (1.0 + norm_b);
        rel_dres = std::sqrt(hc[0]) / (1.0 + norm_c);
        rel_gap = std::fabs(pobj - dobj) / (1.0 + std::fabs(pobj) + std::fabs(dobj));
        const double eps = opt.eps_rel;
        if (opt.verbose > 0)
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
            else if ((double)k >= 0.36 * (double)total) restart = true;
        }
        fpe_prev = fpe;
        if (restart) {
            const double ddx = std::sqrt(hc[2]), ddy = std::sqrt(hr[3]);
            if (ddx > 1e-10 && ddy > 1e-10) w = std::exp(0.5 * std::log(ddy / ddx) + 0.5 * std::log(w));
            ELP_LAUNCH(k_restart, grid1(std::max(n, m)), 256, 0, st, n, x.p, x0.p, xp.p, m, y.p, y0.p, yp.p);
            k = 0; ++restarts; need_fpe0 = true; fpe_prev = -1;
        } else {
            const double wk = (k + 1.0) / (k + 2.0);
            ELP_LAUNCH(k_halpern_finish, grid1(std::max(n, m)), 256, 0, st, n, x.p, xbar.p, x0.p, m, y.p, yp.p, y0.p, wk);
            ++k;
        }
        push_params();
        return false;
    }

    // Farkas-type certificates from the displacement (xp - x0, yp - y0); see oracle/pdlp_ref.py::_certificate
    bool detect_infeasible() {
        double hr[NACC], hc[NACC];
        const double tol = 1e-6;
        // ray vectors: reuse xbar/axbar as scratch is not possible (needed for the Halpern finish) -> gbuf + axp
        if (ray_dx.n == 0) { ray_dx.alloc(n); ray_dy.alloc(std::max(m, 1)); ray_ax.alloc(std::max(m, 1)); ray_g.alloc(n); }
        DevBuf<double>&dxv = ray_dx, &dyv = ray_dy, &axray = ray_ax, &gray = ray_g;
        launch_diff(n, xp.p, x0.p, dxv.p, st);
        launch_diff(m, yp.p, y0.p, dyv.p, st);
        spmv_rows(dxv.p, axray.p);
        ELP_LAUNCH(k_ray_rows, RED_BLOCKS, RED_THREADS, 0, st, m, axray.p, yp.p, y0.p, lc.p, uc.p, dr.p, partials.p);
        fetch_scalars(hr, true);
        spmv_cols(dyv.p, gray.p);
        ELP_LAUNCH(k_ray_cols, RED_BLOCKS, RED_THREADS, 0, st, n, xp.p, x0.p, c.p, l.p, u.p, gray.p, dc.p, partials.p);
        fetch_scalars(hc, false);
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
                if (!dist && opt.time_limit_s > 0 && wall.ms() > opt.time_limit_s * 1e3) break;   // ranks must agree: no wall-clock exit when distributed
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

    void solution(double* x_h, double* y_h, double* obj) {
        DevBuf<double> tmp(std::max(n, std::max(m, 1)));
        if (x_h) {
            ELP_LAUNCH(k_unscale, grid1(n), 256, 0, st, n, xp.p, dc.p, 1.0, tmp.p);
            tmp.download(x_h, n, st);
            ELP_CUDA(cudaStreamSynchronize(st));
            d2h += (int64_t)n * 8;
        }
        if (y_h && m > 0) {
            ELP_LAUNCH(k_unscale, grid1(m), 256, 0, st, m, yp.p, dr.p, maximize ? -1.0 : 1.0, tmp.p);
            tmp.download(y_h, m, st);
            ELP_CUDA(cudaStreamSynchronize(st));
            d2h += (int64_t)m * 8;
        }
        if (obj) *obj = maximize ? -pobj : pobj;
    }

    // Times the two fused iteration kernels in isolation (local part only when distributed: no collective).
    // Leaves the iterate in an arbitrary state: the caller resets afterwards.
    void probe_step(int reps, double* ms_primal, double* ms_dual) {
        cudaEvent_t e0, e1;
        ELP_CUDA(cudaEventCreate(&e0)); ELP_CUDA(cudaEventCreate(&e1));
        float ms = 0;
        PrimalEpi<false> pe{c.p, l.p, u.p, x0.p, x.p, xbar.p, xp.p, params.p, 0};
        auto primal = [&] {
            if (!dist) { launch_spmv(plan_c, n, csc_ptr.p, csc_idx.p, csc_val.p, y.p, pe, st); }
            else {
                launch_spmv(plan_c, n, csc_ptr.p, csc_idx.p, csc_val.p, y.p, StoreEpi{gbuf.p}, st);
                ELP_LAUNCH((apply_epi_kernel<PrimalEpi<false>>), grid1(n), 256, 0, st, n, gbuf.p, pe);
            }
        };
        for (int i = 0; i < 3; ++i) primal();
        ELP_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) primal();
        ELP_CUDA(cudaEventRecord(e1, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        ELP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_primal) *ms_primal = ms / reps;
        for (int i = 0; i < 3; ++i) dual_step<false>(0);
        ELP_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) dual_step<false>(0);
        ELP_CUDA(cudaEventRecord(e1, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        ELP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_dual) *ms_dual = ms / reps;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        reset();
    }

    void probe_spmv(int reps, double* ms_csr, double* ms_csc) {
        cudaEvent_t e0, e1;
        ELP_CUDA(cudaEventCreate(&e0)); ELP_CUDA(cudaEventCreate(&e1));
        float ms = 0;
        for (int i = 0; i < 3; ++i) { spmv_rows(xbar.p, axp.p); }
        ELP_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) spmv_rows(xbar.p, axp.p);
        ELP_CUDA(cudaEventRecord(e1, st));
        ELP_CUDA(cudaStreamSynchronize(st));
        ELP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_csr) *ms_csr = ms / reps;
        for (int i = 0; i < 3; ++i) launch_spmv(plan_c, n, csc_ptr.p, csc_idx.p, csc_val.p, y.p, StoreEpi{gbuf.p}, st);
        ELP_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) launch_spmv(plan_c, n, csc_ptr.p, csc_idx.p, csc_val.p, y.p, StoreEpi{gbuf.p}, st);
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
                  const double* ub, const elp_options& opt, bool dist, elp_stats* stats) {
    WallTimer t;
    auto* p = new Pdlp();
    try {
        p->setup(m, n, row_ptr, col_idx, vals, sense, rhs, c, maximize, lb, ub, opt, dist);
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
    DevBuf<int> ptr(m + 1 + SPMV_PTR_PAD), idx(nnz + SPMV_PAD);
    DevBuf<double> val(nnz + SPMV_PAD), xd(std::max(n, 1)), od(m);
    idx.zero(st); val.zero(st); ptr.zero(st);
    ptr.upload(row_ptr, m + 1, st); idx.upload(col_idx, nnz, st); val.upload(vals, nnz, st); xd.upload(x, n, st);
    // one thread per row: the row sum is formed in index order, bit-identical to a scalar loop
    launch_spmv(plan_spmv(nnz, m, 0, 1), m, ptr.p, idx.p, val.p, xd.p, StoreEpi{od.p}, st);
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
