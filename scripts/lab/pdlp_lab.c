/*
 * pdlp_ref.c — CPU restatement (plain C + OpenMP) of the repo's PDLP loop.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this.
 * It mirrors oracle/pdlp_ref.py statement for statement (r2HPDHG: reflected Halpern PDHG with
 * fixed-point-error restarts, Ruiz + Pock-Chambolle scaling, PDLP relative KKT termination) and is
 * what the CUDA solver in easylp_b200/csrc/pdlp.cu is compared with on small seeded problems
 * (status, objective <= 1e-6 rel, residuals <= 1e-6).  It stands where the reference calls
 * `solve(prob)` (/root/reference/R/class.R:276) for LPs too large for the simplex oracle.
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LE 0
#define GE 1
#define EQ 2

typedef struct {
    int m, n;
    const int *rp, *ci;   /* CSR */
    double *rv;
    int *cp, *ri;         /* CSC */
    double *cv;
} mat_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void spmv_rows(const mat_t *A, const double *x, double *out) {
    int i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < A->m; ++i) {
        double s = 0.0;
        for (int k = A->rp[i]; k < A->rp[i + 1]; ++k) s += A->rv[k] * x[A->ci[k]];
        out[i] = s;
    }
}
static void spmv_cols(const mat_t *A, const double *y, double *out) {
    int j;
#pragma omp parallel for schedule(static)
    for (j = 0; j < A->n; ++j) {
        double s = 0.0;
        for (int k = A->cp[j]; k < A->cp[j + 1]; ++k) s += A->cv[k] * y[A->ri[k]];
        out[j] = s;
    }
}
static double nrm2(const double *v, int n) {
    double s = 0.0;
    int i;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (i = 0; i < n; ++i) s += v[i] * v[i];
    return sqrt(s);
}

/* out[0]=objective (problem's own sense) 1=iterations 2=restarts 3=rel primal res 4=rel dual res 5=rel gap
 * 6=seconds in the iteration loop 7=seconds of setup */
static double envd(const char* n, double d) { const char* e = getenv(n); return e ? atof(e) : d; }
int elpo_pdlp_lab(int m, int n, const int32_t *row_ptr, const int32_t *col_idx, const double *vals,
              const int8_t *sense, const double *rhs, const double *c_in, int maximize,
              const double *lb, const double *ub, double eps, int max_iter, int check_every, int nthreads,
              double *x_out, double *y_out, double *out)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    double t_setup0 = now_s();
    const int64_t nnz = m > 0 ? row_ptr[m] : 0;
    mat_t A;
    A.m = m; A.n = n; A.rp = row_ptr; A.ci = col_idx;
    A.rv = (double *)malloc(sizeof(double) * (nnz + 1));
    memcpy(A.rv, vals, sizeof(double) * nnz);
    A.cp = (int *)calloc((size_t)n + 2, sizeof(int));
    A.ri = (int *)malloc(sizeof(int) * (nnz + 1));
    A.cv = (double *)malloc(sizeof(double) * (nnz + 1));
    int *src = (int *)malloc(sizeof(int) * (nnz + 1));   /* csc slot -> csr slot */
    for (int64_t k = 0; k < nnz; ++k) A.cp[col_idx[k] + 1]++;
    for (int j = 0; j < n; ++j) A.cp[j + 1] += A.cp[j];
    {
        int *cur = (int *)malloc(sizeof(int) * ((size_t)n + 1));
        memcpy(cur, A.cp, sizeof(int) * ((size_t)n + 1));
        for (int i = 0; i < m; ++i)
            for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
                int p = cur[col_idx[k]]++;
                A.ri[p] = i; src[p] = k;
            }
        free(cur);
    }
    double *c = (double *)malloc(sizeof(double) * n), *l = (double *)malloc(sizeof(double) * n),
           *u = (double *)malloc(sizeof(double) * n), *dc = (double *)malloc(sizeof(double) * n);
    double *lc = (double *)malloc(sizeof(double) * (m + 1)), *uc = (double *)malloc(sizeof(double) * (m + 1)),
           *dr = (double *)malloc(sizeof(double) * (m + 1));
    double norm_b = 0, norm_c = 0;
    for (int j = 0; j < n; ++j) { c[j] = maximize ? -c_in[j] : c_in[j]; l[j] = lb[j]; u[j] = ub[j]; dc[j] = 1.0; norm_c += c[j] * c[j]; }
    for (int i = 0; i < m; ++i) {
        lc[i] = sense[i] == LE ? -INFINITY : rhs[i];
        uc[i] = sense[i] == GE ? INFINITY : rhs[i];
        dr[i] = 1.0;
        double a = fmax(isfinite(lc[i]) ? fabs(lc[i]) : 0, isfinite(uc[i]) ? fabs(uc[i]) : 0);
        norm_b += a * a;
    }
    norm_b = sqrt(norm_b); norm_c = sqrt(norm_c);
    /* Ruiz (10) + Pock-Chambolle (alpha = 1) */
    double *rs = (double *)malloc(sizeof(double) * (m + 1)), *cs = (double *)malloc(sizeof(double) * n);
    const int NRUIZ = (int)envd("LAB_RUIZ", 10), USEPC = (int)envd("LAB_PC", 1);
    for (int it = 0; it <= NRUIZ; ++it) {
        const int pc = it == NRUIZ;
        if (pc && !USEPC) break;
        int i, j;
#pragma omp parallel for schedule(static)
        for (i = 0; i < m; ++i) {
            double s = 0;
            for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
                double a = fabs(vals[k]) * dc[col_idx[k]];
                s = pc ? s + a : fmax(s, a);
            }
            rs[i] = s * dr[i];
        }
#pragma omp parallel for schedule(static)
        for (j = 0; j < n; ++j) {
            double s = 0;
            for (int k = A.cp[j]; k < A.cp[j + 1]; ++k) {
                double a = fabs(vals[src[k]]) * dr[A.ri[k]];
                s = pc ? s + a : fmax(s, a);
            }
            cs[j] = s * dc[j];
        }
        for (i = 0; i < m; ++i) if (rs[i] > 0) dr[i] *= 1.0 / sqrt(rs[i]);
        for (j = 0; j < n; ++j) if (cs[j] > 0) dc[j] *= 1.0 / sqrt(cs[j]);
    }
    for (int i = 0; i < m; ++i) {
        for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) A.rv[k] = vals[k] * dr[i] * dc[col_idx[k]];
        lc[i] *= dr[i]; uc[i] *= dr[i];
    }
    for (int j = 0; j < n; ++j) {
        for (int k = A.cp[j]; k < A.cp[j + 1]; ++k) A.cv[k] = A.rv[src[k]];
        c[j] *= dc[j]; l[j] /= dc[j]; u[j] /= dc[j];
    }
    free(rs); free(cs);

    double *x = (double *)calloc(n, sizeof(double)), *x0 = (double *)malloc(sizeof(double) * n),
           *xp = (double *)malloc(sizeof(double) * n), *xbar = (double *)malloc(sizeof(double) * n),
           *g = (double *)malloc(sizeof(double) * n);
    double *y = (double *)calloc(m + 1, sizeof(double)), *y0 = (double *)calloc(m + 1, sizeof(double)),
           *yp = (double *)calloc(m + 1, sizeof(double)), *ax = (double *)calloc(m + 1, sizeof(double)),
           *axp = (double *)calloc(m + 1, sizeof(double));
    /* power iteration */
    double smax = 0.0;
    if (nnz > 0 && m > 0) {
        for (int j = 0; j < n; ++j) {
            uint32_t h = (uint32_t)j * 2654435761u + 12345u;
            h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
            xbar[j] = (double)h / 4294967296.0 - 0.5;
        }
        double nv = nrm2(xbar, n);
        for (int j = 0; j < n; ++j) xbar[j] /= nv;
        double s = 1.0;
        for (int it = 0; it < 60; ++it) {
            spmv_rows(&A, xbar, ax);
            spmv_cols(&A, ax, g);
            double nr = nrm2(g, n);
            if (!(nr > 0)) { s = 0; break; }
            double sn = sqrt(nr);
            for (int j = 0; j < n; ++j) xbar[j] = g[j] / nr;
            int conv = fabs(sn - s) <= 1e-4 * sn;
            s = sn;
            if (conv && it >= 10) break;
        }
        smax = s;
    }
    const double eta = (smax > 0 ? 0.998 / smax : 1.0) * envd("LAB_ETA", 1.0);
    double nb = 0, nc = nrm2(c, n);
    for (int i = 0; i < m; ++i) { double a = fmax(isfinite(lc[i]) ? fabs(lc[i]) : 0, isfinite(uc[i]) ? fabs(uc[i]) : 0); nb += a * a; }
    nb = sqrt(nb);
    double w = (nb > 1e-10 && nc > 1e-10) ? nc / nb : 1.0;
    for (int j = 0; j < n; ++j) { x[j] = fmin(fmax(0.0, l[j]), u[j]); x0[j] = x[j]; xp[j] = x[j]; }

    const double KP = envd("LAB_KP", 0.5), KI = envd("LAB_KI", 0.0), KD = envd("LAB_KD", 0.0);
    const double RHO = envd("LAB_RHO", 1.0), B_SUFF = envd("LAB_BSUFF", 0.2), B_NEC = envd("LAB_BNEC", 0.8), B_ART = envd("LAB_BART", 0.36);
    const double GAPF = envd("LAB_GAPF", 0.25);
    const int OBJTEST = (int)envd("LAB_OBJTEST", 0);
    const double ISMOOTH = envd("LAB_ISMOOTH", 1.0);
    const int WAVG = (int)envd("LAB_WAVG", 0);
    const double PMAX = envd("LAB_PMAX", 1e18);
    const int TRACE = getenv("LAB_TRACE") != NULL;
    double e_int = 0.0, e_prev = 0.0;
    const int ce = check_every > 1 ? check_every : 64;
    const int limit = max_iter > 0 ? max_iter : 2000000;
    int k = 0, total = 0, restarts = 0, status = 7, need_fpe0 = 1;
    double fpe0 = -1, fpe_prev = -1, pobj = 0, dobj = 0, rp = 0, rd = 0, rg = 0;
    double t_setup = now_s() - t_setup0, t0 = now_s();
    while (total < limit) {
        const double tau = eta / w, sig = eta * w;
        const int check = (k % ce) == 0;
        const double wk = (k + 1.0) / (k + 2.0);
        int i, j;
        spmv_cols(&A, y, g);
#pragma omp parallel for schedule(static)
        for (j = 0; j < n; ++j) {
            const double xj = x[j];
            const double xpj = fmin(fmax(xj - tau * (c[j] - g[j]), l[j]), u[j]);
            xbar[j] = 2.0 * xpj - xj;
            if (check) xp[j] = xpj; else x[j] = wk * ((1.0 + RHO) * xpj - RHO * xj) + (1.0 - wk) * x0[j];
        }
        spmv_rows(&A, xbar, ax);
#pragma omp parallel for schedule(static)
        for (i = 0; i < m; ++i) {
            const double yi = y[i], v = yi - sig * ax[i];
            const double lo = v + sig * lc[i], hi = v + sig * uc[i];
            const double ypi = lo > 0 ? lo : (hi < 0 ? hi : 0.0);
            if (check) yp[i] = ypi; else y[i] = wk * ((1.0 + RHO) * ypi - RHO * yi) + (1.0 - wk) * y0[i];
        }
        ++total;
        if (!check) { ++k; continue; }
        /* ---- check: KKT of the candidate T(z) and fixed-point error of z ---- */
        spmv_rows(&A, xp, axp);
        spmv_cols(&A, yp, g);
        double pres2 = 0, dyadx = 0, dy2 = 0, ddy2 = 0, dobj_r = 0, corr_p = 0;
#pragma omp parallel for reduction(+ : pres2, dyadx, dy2, ddy2, dobj_r, corr_p) schedule(static)
        for (i = 0; i < m; ++i) {
            const double a = axp[i];
            const double viol = (a - fmin(fmax(a, lc[i]), uc[i])) / dr[i];
            corr_p += yp[i] * (a - fmin(fmax(a, lc[i]), uc[i]));
            const double dy = yp[i] - y[i], d0 = yp[i] - y0[i];
            pres2 += viol * viol; dyadx += dy * (ax[i] - a); dy2 += dy * dy; ddy2 += d0 * d0;
            dobj_r += yp[i] > 0 ? yp[i] * lc[i] : (yp[i] < 0 ? yp[i] * uc[i] : 0.0);
        }
        double dres2 = 0, dx2 = 0, ddx2 = 0, po = 0, dobj_c = 0, corr_d = 0;
#pragma omp parallel for reduction(+ : dres2, dx2, ddx2, po, dobj_c, corr_d) schedule(static)
        for (j = 0; j < n; ++j) {
            const double r = c[j] - g[j], xpj = xp[j];
            const int at_lo = isfinite(l[j]) && xpj <= l[j], at_hi = isfinite(u[j]) && xpj >= u[j];
            const double rpos = fmax(r, 0.0), rneg = fmin(r, 0.0);
            const double res = ((at_lo ? 0.0 : rpos) + (at_hi ? 0.0 : rneg)) / dc[j];
            corr_d += ((at_lo ? 0.0 : rpos) + (at_hi ? 0.0 : rneg)) * xpj;
            const double dx = xpj - x[j], d0 = xpj - x0[j];
            dres2 += res * res; dx2 += dx * dx; ddx2 += d0 * d0; po += c[j] * xpj;
            dobj_c += (at_lo ? rpos * l[j] : 0.0) + (at_hi ? rneg * u[j] : 0.0);
        }
        const double fpe = sqrt(fmax(dx2 / tau + 2.0 * dyadx + dy2 / sig, 0.0));
        pobj = po; dobj = dobj_r + dobj_c;
        rp = sqrt(pres2) / (1 + norm_b); rd = sqrt(dres2) / (1 + norm_c);
        rg = fabs(pobj - dobj) / (1 + fabs(pobj) + fabs(dobj));
        if (TRACE) fprintf(stderr, "%d %.4e %.4e %.4e %.4e %.4e %.10e\n", total, rp, rd, rg, corr_p / (1 + fabs(pobj) + fabs(dobj)), corr_d / (1 + fabs(pobj) + fabs(dobj)), pobj);
        if (OBJTEST) { const double e = (fabs(pobj - dobj) + fabs(corr_p) + fabs(corr_d)) / (1 + fabs(pobj) + fabs(dobj)); if (rp <= eps && rd <= eps && e <= GAPF * eps) { status = 0; break; } }
        else if (rp <= eps && rd <= eps && rg <= GAPF * eps) { status = 0; break; }   /* gap at eps/4: see pdlp.cu */
        if (need_fpe0) { fpe0 = fpe; need_fpe0 = 0; }
        int restart = 0;
        if (k > 0) {
            if (fpe <= B_SUFF * fpe0) restart = 1;
            else if (fpe <= B_NEC * fpe0 && fpe_prev >= 0 && fpe > fpe_prev) restart = 1;
            else if ((double)k >= B_ART * (double)total || (double)k >= PMAX) restart = 1;
        }
        fpe_prev = fpe;
        if (restart) {
            const double ddx = sqrt(ddx2), ddy = sqrt(ddy2);
            if (ddx > 1e-10 && ddy > 1e-10) {
                const double e = log(w * ddx / ddy);      /* >0: primal moves too much relative to dual */
                e_int = KI * e_int + e;
                w = exp(log(w) - (KP * e + (KI > 0 ? KI * (e_int - e) : 0.0) + KD * (e - e_prev)));
                e_prev = e;
                (void)WAVG;
            }
            for (j = 0; j < n; ++j) { x[j] = xp[j]; x0[j] = xp[j]; }
            for (i = 0; i < m; ++i) { y[i] = yp[i]; y0[i] = yp[i]; }
            k = 0; ++restarts; need_fpe0 = 1; fpe_prev = -1;
        } else {
            for (j = 0; j < n; ++j) x[j] = wk * ((1.0 + RHO) * xp[j] - RHO * x[j]) + (1.0 - wk) * x0[j];
            for (i = 0; i < m; ++i) y[i] = wk * ((1.0 + RHO) * yp[i] - RHO * y[i]) + (1.0 - wk) * y0[i];
            ++k;
        }
    }
    double t_loop = now_s() - t0;
    if (x_out) for (int j = 0; j < n; ++j) x_out[j] = xp[j] * dc[j];
    if (y_out) for (int i = 0; i < m; ++i) y_out[i] = (maximize ? -1.0 : 1.0) * yp[i] * dr[i];
    if (out) {
        out[0] = maximize ? -pobj : pobj; out[1] = total; out[2] = restarts; out[3] = rp; out[4] = rd; out[5] = rg;
        out[6] = t_loop; out[7] = t_setup;
    }
    free(A.rv); free(A.cp); free(A.ri); free(A.cv); free(src); free(c); free(l); free(u); free(dc); free(lc); free(uc);
    free(dr); free(x); free(x0); free(xp); free(xbar); free(g); free(y); free(y0); free(yp); free(ax); free(axp);
    return status;
}
