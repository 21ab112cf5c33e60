"""CPU restatement (numpy/scipy.sparse) of the repo's PDLP loop — TEST INFRASTRUCTURE ONLY.

Nothing in the product path (easylp_b200/, include/, csrc/) may import this file.  It exists so that
tests can compare the CUDA r2HPDHG iteration against an independent, readable CPU statement of the
same published algorithm, and so that bench.py's `cpu_baseline` leg has something to time.

Algorithm: restarted, reflected Halpern PDHG for LP ("r2HPDHG", Lu & Yang 2024; the scheme used by
cuPDLPx), on the two-sided form
        min c'x   s.t.  lc <= A x <= uc ,  l <= x <= u
with Ruiz + Pock-Chambolle diagonal preconditioning, constant step size eta = 0.998/||A||_2,
primal weight w updated at restarts, fixed-point-error restarts (0.2 / 0.8 / 0.36 rules) and
PDLP-style relative KKT termination (primal residual, dual residual, gap <= eps).

Parity: this solver replaces the reference's `solve(prob)` call at /root/reference/R/class.R:276
(lp_solve, absent from the image) for *large* LPs.  Parity is anchored on status + objective
(<=1e-6 rel) + residuals (<=1e-6 rel), not on iterates.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

LE, GE, EQ = 0, 1, 2
INF = np.inf

STATUS_OPTIMAL = 0
STATUS_INFEASIBLE = 2
STATUS_UNBOUNDED = 3
STATUS_NUMERICAL = 5
STATUS_TIMEOUT = 7   # iteration limit maps onto lp_solve's TIMEOUT (R/class.R:288)


def row_bounds(sense, rhs):
    lc = np.where(sense == LE, -INF, rhs)
    uc = np.where(sense == GE, INF, rhs)
    return lc, uc


def ruiz_pc_scaling(A: sp.csr_matrix, ruiz_iters=10, pc_alpha=1.0):
    """Returns (dr, dc) with scaled A = diag(dr) A diag(dc)."""
    m, n = A.shape
    dr = np.ones(m)
    dc = np.ones(n)
    if m == 0 or n == 0 or A.nnz == 0:
        return dr, dc
    As = A.copy().tocsr()
    absA = abs(As)
    for _ in range(ruiz_iters):
        rmax = absA.max(axis=1).toarray().ravel()
        cmax = absA.max(axis=0).toarray().ravel()
        r = 1.0 / np.sqrt(np.where(rmax > 0, rmax, 1.0))
        c = 1.0 / np.sqrt(np.where(cmax > 0, cmax, 1.0))
        absA = sp.diags(r) @ absA @ sp.diags(c)
        dr *= r
        dc *= c
    # Pock-Chambolle alpha=1: row scale 1/sqrt(row L1), col scale 1/sqrt(col L1)
    rs = np.asarray(absA.sum(axis=1)).ravel()
    cs = np.asarray(absA.sum(axis=0)).ravel()
    r = 1.0 / np.sqrt(np.where(rs > 0, rs, 1.0))
    c = 1.0 / np.sqrt(np.where(cs > 0, cs, 1.0))
    dr *= r
    dc *= c
    return dr, dc


def power_sigma_max(A, AT, iters=40, seed=1):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(A.shape[1])
    v /= np.linalg.norm(v)
    s = 1.0
    for _ in range(iters):
        u = A @ v
        v = AT @ u
        s = np.linalg.norm(v)
        if s == 0:
            return 0.0
        v /= s
    return float(np.sqrt(s))


def dual_prox(v, sigma, lc, uc):
    """argmax_y  -(y - v)^2/(2 sigma) + y+ lc - y- uc  evaluated at v = y - sigma*A xbar."""
    lo = v + sigma * lc          # -inf where lc = -inf
    hi = v + sigma * uc          # +inf where uc = +inf
    return np.where(lo > 0, lo, np.where(hi < 0, hi, 0.0))


class Result(dict):
    __getattr__ = dict.get


def kkt(A, AT, c, lc, uc, l, u, x, y, ax=None, aty=None, at_lo=None, at_hi=None):
    """Unscaled-space KKT quantities for (x, y).
    Reduced costs r = c - A'y are split PDLP-style: r_j > 0 with x_j sitting AT a finite lower bound
    (or r_j < 0 at a finite upper bound) enters the dual objective through that bound; any other
    part of r_j is a dual residual ("primal gradient on finite bounds treated as residual")."""
    if ax is None:
        ax = A @ x
    if aty is None:
        aty = AT @ y
    if at_lo is None:
        at_lo = np.isfinite(l) & (x <= l)
    if at_hi is None:
        at_hi = np.isfinite(u) & (x >= u)
    pres = ax - np.clip(ax, lc, uc)
    r = c - aty
    rpos = np.maximum(r, 0.0)
    rneg = np.minimum(r, 0.0)
    dres_v = np.where(at_lo, 0.0, rpos) + np.where(at_hi, 0.0, rneg)
    pobj = float(c @ x)
    lfin = np.where(at_lo, l, 0.0)
    ufin = np.where(at_hi, u, 0.0)
    lcf = np.where(np.isfinite(lc), lc, 0.0)
    ucf = np.where(np.isfinite(uc), uc, 0.0)
    ypos = np.maximum(y, 0.0)
    yneg = np.minimum(y, 0.0)
    dobj = float(ypos @ lcf + yneg @ ucf + rpos @ lfin + rneg @ ufin)
    return dict(pres=float(np.linalg.norm(pres)), dres=float(np.linalg.norm(dres_v)),
                pobj=pobj, dobj=dobj, gap=abs(pobj - dobj))


def solve(m, n, row_ptr, col_idx, vals, sense, rhs, c, lb, ub, maximize=False,
          eps=1e-6, max_iter=200_000, check_every=64, verbose=False, ruiz_iters=10,
          kp=0.85, ki=0.0, kd=0.0, i_smooth=0.3, history=None):
    A0 = sp.csr_matrix((vals, col_idx, row_ptr), shape=(m, n))
    c0 = -np.asarray(c, dtype=float) if maximize else np.asarray(c, dtype=float)
    lc0, uc0 = row_bounds(np.asarray(sense), np.asarray(rhs, dtype=float))
    l0 = np.asarray(lb, dtype=float)
    u0 = np.asarray(ub, dtype=float)

    dr, dc = ruiz_pc_scaling(A0, ruiz_iters)
    A = (sp.diags(dr) @ A0 @ sp.diags(dc)).tocsr()
    AT = A.T.tocsr()
    A0T = A0.T.tocsr()
    cs = c0 * dc
    lcs, ucs = lc0 * dr, uc0 * dr
    ls, us = l0 / dc, u0 / dc

    bfin = np.where(np.isfinite(lc0), lc0, np.where(np.isfinite(uc0), uc0, 0.0))
    # PDLP's combined rhs norm: use the finite bound of larger magnitude per row
    bboth = np.maximum(np.abs(np.where(np.isfinite(lc0), lc0, 0.0)), np.abs(np.where(np.isfinite(uc0), uc0, 0.0)))
    norm_b = float(np.linalg.norm(bboth))
    norm_c = float(np.linalg.norm(c0))

    smax = power_sigma_max(A, AT)
    eta = 0.998 / smax if smax > 0 else 1.0
    bs = np.maximum(np.abs(np.where(np.isfinite(lcs), lcs, 0.0)), np.abs(np.where(np.isfinite(ucs), ucs, 0.0)))
    nb, nc = np.linalg.norm(bs), np.linalg.norm(cs)
    w = nc / nb if nb > 1e-10 and nc > 1e-10 else 1.0

    x = np.clip(np.zeros(n), ls, us)
    y = np.zeros(m)
    x0, y0 = x.copy(), y.copy()
    k = 0                      # iterations since restart
    total = 0
    fpe0 = None                # fixed-point error at restart anchor
    fpe_prev = None
    e_int = 0.0
    e_prev = 0.0
    status = STATUS_TIMEOUT
    best = None
    n_restart = 0
    while total < max_iter:
        tau, sig = eta / w, eta * w
        aty = AT @ y
        xp = np.clip(x - tau * (cs - aty), ls, us)
        xbar = 2 * xp - x
        axbar = A @ xbar
        yp = dual_prox(y - sig * axbar, sig, lcs, ucs)
        yref = 2 * yp - y
        do_check = (total % check_every == check_every - 1) or total == 0
        if do_check:
            # fixed-point error  || z - T(z) ||_P
            dx, dy = xp - x, yp - y
            adx = A @ dx
            fpe = np.sqrt(max(dx @ dx / tau + 2 * (dy @ adx) + dy @ dy / sig, 0.0))
            if fpe0 is None:
                fpe0 = fpe
            # KKT on the candidate T(z), unscaled
            xu, yu = xp * dc, yp * dr
            q = kkt(A0, A0T, c0, lc0, uc0, l0, u0, xu, yu, at_lo=np.isfinite(ls) & (xp <= ls), at_hi=np.isfinite(us) & (xp >= us))
            if history is not None:
                history.append((total + 1, q["pres"], q["dres"], q["gap"], fpe, w))
            if verbose:
                print(f"it {total+1:7d} pres {q['pres']:.3e} dres {q['dres']:.3e} gap {q['gap']:.3e} "
                      f"pobj {q['pobj']:.9e} fpe {fpe:.3e} w {w:.3e} restarts {n_restart}")
            ok = (q["pres"] <= eps * (1 + norm_b) and q["dres"] <= eps * (1 + norm_c)
                  and q["gap"] <= 0.25 * eps * (1 + abs(q["pobj"]) + abs(q["dobj"])))   # eps/4: see pdlp.cu
            if ok:
                status = STATUS_OPTIMAL
                best = (xu, yu, q)
                total += 1
                break
            # infeasibility / unboundedness certificates from the displacement
            cert = _certificate(A0, A0T, c0, lc0, uc0, l0, u0, (xp - x0) * dc, (yp - y0) * dr, eps)
            if cert is not None and k > 0:
                status = cert
                best = (xu, yu, q)
                total += 1
                break
            restart = False
            if k > 0:
                if fpe <= 0.2 * fpe0:
                    restart = True
                elif fpe <= 0.8 * fpe0 and fpe_prev is not None and fpe > fpe_prev:
                    restart = True
                elif k >= 0.36 * total:
                    restart = True
            fpe_prev = fpe
            if restart:
                ddx = np.linalg.norm(xp - x0)
                ddy = np.linalg.norm(yp - y0)
                if ddx > 1e-10 and ddy > 1e-10:
                    e = np.log(w * ddx / ddy)
                    e_int = i_smooth * e_int + e   # (leaky) integral term
                    w = w * np.exp(-(kp * e + ki * e_int + kd * (e - e_prev)))
                    e_prev = e
                x, y = xp.copy(), yp.copy()
                x0, y0 = x.copy(), y.copy()
                k = 0
                fpe0 = None
                fpe_prev = None
                n_restart += 1
                total += 1
                # fpe0 for the new epoch is computed on its first iteration
                tau, sig = eta / w, eta * w
                aty = AT @ y
                xq = np.clip(x - tau * (cs - aty), ls, us)
                yq = dual_prox(y - sig * (A @ (2 * xq - x)), sig, lcs, ucs)
                dx, dy = xq - x, yq - y
                fpe0 = np.sqrt(max(dx @ dx / tau + 2 * (dy @ (A @ dx)) + dy @ dy / sig, 0.0))
                continue
        wk = (k + 1.0) / (k + 2.0)
        x = wk * xbar + (1 - wk) * x0
        y = wk * yref + (1 - wk) * y0
        k += 1
        total += 1
    if best is None:
        xu, yu = x * dc, y * dr
        best = (xu, yu, kkt(A0, A0T, c0, lc0, uc0, l0, u0, xu, yu))
    xu, yu, q = best
    obj = q["pobj"]
    if maximize:
        obj = -obj
    return Result(status=status, x=xu, y=yu, obj=obj, iters=total, restarts=n_restart,
                  pres=q["pres"] / (1 + norm_b), dres=q["dres"] / (1 + norm_c),
                  gap=q["gap"] / (1 + abs(q["pobj"]) + abs(q["dobj"])))


def _certificate(A, AT, c, lc, uc, l, u, dx, dy, eps):
    """Farkas-type certificates from the displacement (dx, dy) = z - z_anchor.
    Returns STATUS_INFEASIBLE / STATUS_UNBOUNDED / None."""
    tol = 1e-6
    # primal infeasibility: dual ray dy with A'dy ~ 0 w.r.t. free directions and positive dual ray objective
    ny = np.linalg.norm(dy, np.inf)
    if ny > 1e-12:
        yr = dy / ny
        ypos, yneg = np.maximum(yr, 0), np.minimum(yr, 0)
        # rays must respect row-bound finiteness
        bad = (ypos @ np.where(np.isfinite(lc), 0.0, 1.0)) + (-yneg @ np.where(np.isfinite(uc), 0.0, 1.0))
        r = -(AT @ yr)
        rpos, rneg = np.maximum(r, 0), np.minimum(r, 0)
        res = np.where(np.isfinite(l), 0.0, rpos) + np.where(np.isfinite(u), 0.0, rneg)
        lf, uf = np.where(np.isfinite(l), l, 0.0), np.where(np.isfinite(u), u, 0.0)
        lcf, ucf = np.where(np.isfinite(lc), lc, 0.0), np.where(np.isfinite(uc), uc, 0.0)
        dobj = ypos @ lcf + yneg @ ucf + rpos @ lf + rneg @ uf
        if dobj > 0 and (np.linalg.norm(res) + bad) / dobj <= tol:
            return STATUS_INFEASIBLE
    nx = np.linalg.norm(dx, np.inf)
    if nx > 1e-12:
        xr = dx / nx
        cobj = c @ xr
        if cobj < 0:
            axr = A @ xr
            viol = np.where(np.isfinite(lc), np.minimum(axr, 0.0), 0.0) + np.where(np.isfinite(uc), np.maximum(axr, 0.0), 0.0)
            vb = np.where(np.isfinite(l), np.minimum(xr, 0.0), 0.0) + np.where(np.isfinite(u), np.maximum(xr, 0.0), 0.0)
            if (np.linalg.norm(viol) + np.linalg.norm(vb)) / (-cobj) <= tol:
                return STATUS_UNBOUNDED
    return None
