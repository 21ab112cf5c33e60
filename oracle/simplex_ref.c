/*
 * simplex_ref.c — CPU oracle for the small-LP solve path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this file's shared object.  The product (easylp_b200/, include/) never links or calls it.
 *
 * What it restates: the call `status <- solve(prob)` at /root/reference/R/class.R:276 and the model
 * load in front of it (R/class.R:260-274) — i.e. lp_solve 5.5's job as used by the reference:
 *   min/max c'x  s.t.  rows  a_i'x {<=,>=,=} rhs_i ,  lb <= x <= ub  (default bounds -Inf..+Inf,
 *   R/class.R:86), returning lp_solve's integer status (0 optimal, 2 infeasible, 3 unbounded),
 *   the objective and the variable values.
 * lp_solve itself (lpSolveAPI, CRAN, un-pinned in /root/reference/DESCRIPTION:18-21) is NOT in the
 * image, so this is a restatement of its published algorithm class — the bounded-variable primal
 * simplex method with a composite phase 1 (minimise the sum of infeasibilities), Dantzig pricing,
 * a two-pass ratio test and Bland's rule against cycling — not a port of lp_solve's code.
 * An LP's status and optimal objective do not depend on the pivoting rule, so parity on
 * status/objective is what this oracle pins (goldens G1,G2,G3,G5 in tests/golden/, plus HiGHS
 * cross-checks in tests/test_oracle_simplex.py).
 *
 * Dense tableau, one LP per call; `elpo_simplex_batch` loops over a batch (OpenMP over LPs).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ST_BASIC 0
#define ST_LOWER 1
#define ST_UPPER 2
#define ST_FREE 3

#define TOL_PRIMAL 1e-9
#define TOL_DUAL 1e-9
#define TOL_PIVOT 1e-9

typedef struct {
    int m, n, N;
    double *T;      /* m x N tableau = B^-1 [A | I] */
    double *beta;   /* m basic values */
    double *lo, *hi, *cost, *xn, *d;   /* N */
    int *basis;     /* m */
    int8_t *state;  /* N */
    double *cb;     /* m scratch */
} lp_t;

static double ptol(double bound) { return TOL_PRIMAL * fmax(1.0, fabs(bound)); }

/* returns lp_solve status; fills objval, x[n], y[m] (y may be NULL), pivots */
int elpo_simplex(int m, int n, const double *A, const double *b, const int8_t *sense,
                 const double *c, int maximize, const double *lb, const double *ub,
                 int max_pivots, double *objval, double *x, double *y, int *pivots_out)
{
    lp_t L;
    int N = n + m, i, j, status = 7, pivots = 0, degenerate_run = 0, bland = 0;
    L.m = m; L.n = n; L.N = N;
    L.T = (double *)calloc((size_t)(m > 0 ? m : 1) * N, sizeof(double));
    L.beta = (double *)calloc(m + 1, sizeof(double));
    L.lo = (double *)calloc(N + 1, sizeof(double));
    L.hi = (double *)calloc(N + 1, sizeof(double));
    L.cost = (double *)calloc(N + 1, sizeof(double));
    L.xn = (double *)calloc(N + 1, sizeof(double));
    L.d = (double *)calloc(N + 1, sizeof(double));
    L.basis = (int *)calloc(m + 1, sizeof(int));
    L.state = (int8_t *)calloc(N + 1, sizeof(int8_t));
    L.cb = (double *)calloc(m + 1, sizeof(double));
    if (max_pivots <= 0) max_pivots = 50 * (m + n) + 1000;

    int bad_bounds = 0;
    for (j = 0; j < n; ++j) {
        L.lo[j] = lb ? lb[j] : 0.0;
        L.hi[j] = ub ? ub[j] : INFINITY;
        L.cost[j] = maximize ? -c[j] : c[j];
        if (L.lo[j] > L.hi[j]) bad_bounds = 1;
        if (isfinite(L.lo[j])) { L.state[j] = ST_LOWER; L.xn[j] = L.lo[j]; }
        else if (isfinite(L.hi[j])) { L.state[j] = ST_UPPER; L.xn[j] = L.hi[j]; }
        else { L.state[j] = ST_FREE; L.xn[j] = 0.0; }
    }
    for (i = 0; i < m; ++i) {
        int s = sense ? sense[i] : 0;
        j = n + i;
        L.lo[j] = (s == 1) ? -INFINITY : 0.0;
        L.hi[j] = (s == 0) ? INFINITY : 0.0;
        L.cost[j] = 0.0;
        L.state[j] = ST_BASIC;
        L.basis[i] = j;
        double r = b[i];
        for (int k = 0; k < n; ++k) {
            L.T[(size_t)i * N + k] = A[(size_t)i * n + k];
            r -= A[(size_t)i * n + k] * L.xn[k];
        }
        L.T[(size_t)i * N + n + i] = 1.0;
        L.beta[i] = r;
    }
    if (bad_bounds) { status = 2; goto finish; }

    for (;;) {
        /* phase detection: gradient of the sum of infeasibilities w.r.t. the basic values */
        double w = 0.0;
        for (i = 0; i < m; ++i) {
            int k = L.basis[i];
            double g = 0.0;
            if (L.beta[i] < L.lo[k] - ptol(L.lo[k])) { g = -1.0; w += L.lo[k] - L.beta[i]; }
            else if (L.beta[i] > L.hi[k] + ptol(L.hi[k])) { g = 1.0; w += L.beta[i] - L.hi[k]; }
            L.cb[i] = g;
        }
        int phase1 = w > 0.0;
        if (!phase1) for (i = 0; i < m; ++i) L.cb[i] = L.cost[L.basis[i]];
        /* reduced costs d_j = cost_j - cb' T[:,j] */
        for (j = 0; j < N; ++j) {
            double s = phase1 ? 0.0 : L.cost[j];
            for (i = 0; i < m; ++i) s -= L.cb[i] * L.T[(size_t)i * N + j];
            L.d[j] = s;
        }
        /* pricing */
        int q = -1; double best = 0.0; int dir = 0;
        for (j = 0; j < N; ++j) {
            int st = L.state[j];
            if (st == ST_BASIC) continue;
            if (!(L.lo[j] < L.hi[j])) continue;            /* fixed */
            double dj = L.d[j]; int dd = 0;
            if ((st == ST_LOWER || st == ST_FREE) && dj < -TOL_DUAL) dd = 1;
            else if ((st == ST_UPPER || st == ST_FREE) && dj > TOL_DUAL) dd = -1;
            if (!dd) continue;
            if (bland) { q = j; dir = dd; break; }
            if (fabs(dj) > best) { best = fabs(dj); q = j; dir = dd; }
        }
        if (q < 0) { status = phase1 ? 2 : 0; break; }
        if (pivots >= max_pivots) { status = 7; break; }

        /* ratio test, pass 1: minimum step */
        double tmin = INFINITY;
        double tflip = L.hi[q] - L.lo[q];                 /* inf if either bound is infinite */
        if (L.state[q] == ST_FREE) tflip = INFINITY;
        for (i = 0; i < m; ++i) {
            double a = dir * L.T[(size_t)i * N + q];
            if (fabs(a) <= TOL_PIVOT) continue;
            int k = L.basis[i]; double bi = L.beta[i], t = INFINITY;
            if (a > 0) {                                   /* beta_i decreases */
                if (bi > L.hi[k] + ptol(L.hi[k])) t = (bi - L.hi[k]) / a;
                else if (bi >= L.lo[k] - ptol(L.lo[k])) { if (isfinite(L.lo[k])) t = fmax(bi - L.lo[k], 0.0) / a; }
            } else {                                       /* beta_i increases */
                if (bi < L.lo[k] - ptol(L.lo[k])) t = (bi - L.lo[k]) / a;
                else if (bi <= L.hi[k] + ptol(L.hi[k])) { if (isfinite(L.hi[k])) t = fmin(bi - L.hi[k], 0.0) / a; }
            }
            if (t < tmin) tmin = t;
        }
        if (tflip <= tmin) {
            if (!isfinite(tflip)) {                        /* no blocking row, no opposite bound */
                if (phase1) { status = 5; break; }
                status = 3;
                L.xn[q] = dir > 0 ? INFINITY : -INFINITY;
                break;
            }
            /* bound flip */
            for (i = 0; i < m; ++i) L.beta[i] -= dir * tflip * L.T[(size_t)i * N + q];
            if (dir > 0) { L.state[q] = ST_UPPER; L.xn[q] = L.hi[q]; }
            else { L.state[q] = ST_LOWER; L.xn[q] = L.lo[q]; }
            ++pivots;
            degenerate_run = 0; bland = 0;
            continue;
        }
        /* pass 2: among rows within a tiny window of tmin take the largest pivot
           (Bland mode: the smallest basic index) */
        int r = -1; double amax = 0.0; int to_upper = 0;
        double window = tmin + 1e-12 * fmax(1.0, fabs(tmin));
        for (i = 0; i < m; ++i) {
            double a = dir * L.T[(size_t)i * N + q];
            if (fabs(a) <= TOL_PIVOT) continue;
            int k = L.basis[i]; double bi = L.beta[i], t = INFINITY; int up = 0;
            if (a > 0) {
                if (bi > L.hi[k] + ptol(L.hi[k])) { t = (bi - L.hi[k]) / a; up = 1; }
                else if (bi >= L.lo[k] - ptol(L.lo[k])) { if (isfinite(L.lo[k])) { t = fmax(bi - L.lo[k], 0.0) / a; up = 0; } }
            } else {
                if (bi < L.lo[k] - ptol(L.lo[k])) { t = (bi - L.lo[k]) / a; up = 0; }
                else if (bi <= L.hi[k] + ptol(L.hi[k])) { if (isfinite(L.hi[k])) { t = fmin(bi - L.hi[k], 0.0) / a; up = 1; } }
            }
            if (t > window) continue;
            if (bland) { if (r < 0 || L.basis[i] < L.basis[r]) { r = i; to_upper = up; } }
            else if (fabs(a) > amax) { amax = fabs(a); r = i; to_upper = up; }
        }
        if (r < 0) { status = 5; break; }
        /* step */
        double t = tmin;
        double xq = L.xn[q] + dir * t;
        for (i = 0; i < m; ++i) L.beta[i] -= dir * t * L.T[(size_t)i * N + q];
        int kleave = L.basis[r];
        if (to_upper) { L.state[kleave] = ST_UPPER; L.xn[kleave] = L.hi[kleave]; }
        else { L.state[kleave] = ST_LOWER; L.xn[kleave] = L.lo[kleave]; }
        L.beta[r] = xq;
        L.basis[r] = q;
        L.state[q] = ST_BASIC;
        /* rank-1 tableau update */
        double piv = L.T[(size_t)r * N + q];
        double *Tr = L.T + (size_t)r * N;
        for (j = 0; j < N; ++j) Tr[j] /= piv;
        Tr[q] = 1.0;
        for (i = 0; i < m; ++i) {
            if (i == r) continue;
            double f = L.T[(size_t)i * N + q];
            if (f == 0.0) continue;
            double *Ti = L.T + (size_t)i * N;
            for (j = 0; j < N; ++j) Ti[j] -= f * Tr[j];
            Ti[q] = 0.0;
        }
        ++pivots;
        if (t <= 1e-12) { if (++degenerate_run > 30) bland = 1; }
        else { degenerate_run = 0; bland = 0; }
    }

finish:
    for (j = 0; j < N; ++j) if (L.state[j] != ST_BASIC && j < n) x[j] = L.xn[j];
    for (i = 0; i < m; ++i) if (L.basis[i] < n) x[L.basis[i]] = L.beta[i];
    {
        double obj = 0.0;
        if (status == 3) obj = -INFINITY;
        else for (j = 0; j < n; ++j) obj += L.cost[j] * x[j];
        *objval = maximize ? -obj : obj;
    }
    if (y) {
        for (i = 0; i < m; ++i) {
            double s = 0.0;
            for (int k = 0; k < m; ++k) s += L.cost[L.basis[k]] * L.T[(size_t)k * N + n + i];
            y[i] = maximize ? -s : s;
        }
    }
    if (pivots_out) *pivots_out = pivots;
    free(L.T); free(L.beta); free(L.lo); free(L.hi); free(L.cost); free(L.xn); free(L.d);
    free(L.basis); free(L.state); free(L.cb);
    return status;
}

/* batch: A[B][m][n], b[B][m], c[B][n], lb/ub [B][n] or NULL, sense [B][m] or NULL */
void elpo_simplex_batch(int64_t B, int m, int n, const double *A, const double *b, const double *c,
                        const double *lb, const double *ub, const int8_t *sense, int maximize,
                        int32_t *status, double *obj, double *x, int32_t *pivots, int nthreads)
{
    int64_t k;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (k = 0; k < B; ++k) {
        int piv = 0;
        status[k] = elpo_simplex(m, n, A + k * m * n, b + k * m, sense ? sense + k * m : NULL, c + k * n,
                                 maximize, lb ? lb + k * n : NULL, ub ? ub + k * n : NULL, 0,
                                 obj + k, x + k * n, NULL, &piv);
        if (pivots) pivots[k] = piv;
    }
}

/* CSR front end used for the single-LP goldens */
int elpo_simplex_csr(int m, int n, const int32_t *row_ptr, const int32_t *col_idx, const double *vals,
                     const int8_t *sense, const double *rhs, const double *c, int maximize,
                     const double *lb, const double *ub, double *objval, double *x, double *y, int *pivots)
{
    double *A = (double *)calloc((size_t)(m > 0 ? m : 1) * (n > 0 ? n : 1), sizeof(double));
    for (int i = 0; i < m; ++i)
        for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) A[(size_t)i * n + col_idx[k]] += vals[k];
    int st = elpo_simplex(m, n, A, rhs, sense, c, maximize, lb, ub, 0, objval, x, y, pivots);
    free(A);
    return st;
}
