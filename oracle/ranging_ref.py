"""CPU restatement of easylp_b200/csrc/sensitivity.cu: textbook basis-invariance ranging.  TEST INFRASTRUCTURE ONLY.

Stands where the reference calls lpSolveAPI::get.sensitivity.obj / get.sensitivity.rhs
(/root/reference/R/class.R:624,641).  lp_solve is not in this image and the reference's tests print no sensitivity
values, so the pin is HiGHS' own ranging (scipy's bundled HiGHS core) on non-degenerate models: objective ranges of the
basic columns and right-hand-side ranges of the binding rows must agree (tests/test_sensitivity.py).  What lp_solve
reports at degenerate vertices or for non-basic columns is NOT pinned ("parity unpinned" for those)."""
from __future__ import annotations

import numpy as np


def highs(p):
    """Optimal basis and HiGHS' ranging of a problem dict (keys of oracle/gen.py).  Returns a dict or None when HiGHS'
    python core is unavailable."""
    try:
        from scipy.optimize._highspy import _core as H
    except Exception:
        return None
    m, n = int(p["m"]), int(p["n"])
    h = H._Highs()
    h.setOptionValue("output_flag", False)
    lp = H.HighsLp()
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = np.asarray(p["c"], float)
    lp.col_lower_ = np.broadcast_to(np.asarray(p["lb"], float), (n,)).copy()
    lp.col_upper_ = np.broadcast_to(np.asarray(p["ub"], float), (n,)).copy()
    s = np.asarray(p["sense"])
    lp.row_lower_ = np.where(s == 0, -np.inf, p["rhs"]).astype(float)
    lp.row_upper_ = np.where(s == 1, np.inf, p["rhs"]).astype(float)
    lp.sense_ = H.ObjSense.kMaximize if p.get("maximize") else H.ObjSense.kMinimize
    lp.a_matrix_.format_ = H.MatrixFormat.kRowwise
    lp.a_matrix_.start_ = np.asarray(p["row_ptr"], np.int32)
    lp.a_matrix_.index_ = np.asarray(p["col_idx"], np.int32)
    lp.a_matrix_.value_ = np.asarray(p["vals"], float)
    h.passModel(lp)
    h.run()
    if h.getModelStatus() != H.HighsModelStatus.kOptimal:
        return None
    sol, bas = h.getSolution(), h.getBasis()
    _, r = h.getRanging()
    code = {H.HighsBasisStatus.kBasic: 0, H.HighsBasisStatus.kLower: -1, H.HighsBasisStatus.kUpper: 1,
            H.HighsBasisStatus.kZero: 2, H.HighsBasisStatus.kNonbasic: 2}
    return dict(x=np.array(sol.col_value), col_status=np.array([code[v] for v in bas.col_status]),
                row_status=np.array([code[v] for v in bas.row_status]),
                cost_dn=np.array(r.col_cost_dn.value_)[:n], cost_up=np.array(r.col_cost_up.value_)[:n],
                row_dn=np.array(r.row_bound_dn.value_)[:m], row_up=np.array(r.row_bound_up.value_)[:m])


def ranging(p, x, col_basic, row_basic):
    """Textbook ranging for the basis {columns with col_basic} + {slacks of rows with row_basic}.
    Returns (obj_from, obj_till, rhs_from, rhs_till) in the problem's own sense."""
    m, n = int(p["m"]), int(p["n"])
    A = np.zeros((m, n))
    rp = np.asarray(p["row_ptr"])
    for i in range(m):
        A[i, p["col_idx"][rp[i]:rp[i + 1]]] = p["vals"][rp[i]:rp[i + 1]]
    AI = np.hstack([A, np.eye(m)])
    sgn = -1.0 if p.get("maximize") else 1.0
    cm = np.concatenate([sgn * np.asarray(p["c"], float), np.zeros(m)])
    s = np.asarray(p["sense"])
    lo = np.concatenate([np.broadcast_to(p["lb"], (n,)), np.where(s == 1, -np.inf, 0.0)]).astype(float)
    up = np.concatenate([np.broadcast_to(p["ub"], (n,)), np.where(s == 0, np.inf, 0.0)]).astype(float)
    xv = np.concatenate([x, np.asarray(p["rhs"], float) - A @ x])
    basic = np.concatenate([np.asarray(col_basic, bool), np.asarray(row_basic, bool)])
    B = np.nonzero(basic)[0]
    assert B.size == m, "not a basis"
    Binv = np.linalg.inv(AI[:, B])
    pi = cm[B] @ Binv
    d = cm - pi @ AI
    tol = 1e-9
    at = np.zeros(n + m, int)
    for q in np.nonzero(~basic)[0]:
        if np.isfinite(lo[q]) and np.isfinite(up[q]) and up[q] - lo[q] <= tol:
            at[q] = 0
        elif np.isfinite(lo[q]) and abs(xv[q] - lo[q]) <= tol * (1 + abs(lo[q])):
            at[q] = -1
        elif np.isfinite(up[q]) and abs(xv[q] - up[q]) <= tol * (1 + abs(up[q])):
            at[q] = 1
    of, ot = np.zeros(n), np.zeros(n)
    alpha = Binv @ AI                     # row k: B^-1 a_q over all q
    for j in range(n):
        dmin, dmax = -np.inf, np.inf
        if not basic[j]:
            if at[j] < 0:
                dmin = -d[j]
            elif at[j] > 0:
                dmax = -d[j]
            elif not (np.isfinite(lo[j]) and np.isfinite(up[j]) and up[j] - lo[j] <= tol):
                dmin = dmax = 0.0
        else:
            k = int(np.nonzero(B == j)[0][0])
            for q in np.nonzero(~basic)[0]:
                if at[q] == 0 or abs(alpha[k, q]) <= 1e-12:
                    continue
                ratio = d[q] / alpha[k, q]
                if (at[q] < 0 and alpha[k, q] > 0) or (at[q] > 0 and alpha[k, q] < 0):
                    dmax = min(dmax, ratio)
                else:
                    dmin = max(dmin, ratio)
        a, b = cm[j] + dmin, cm[j] + dmax
        of[j], ot[j] = (-b, -a) if p.get("maximize") else (a, b)
    rf, rt = np.zeros(m), np.zeros(m)
    for i in range(m):
        dmin, dmax = -np.inf, np.inf
        for k, q in enumerate(B):
            g = Binv[k, i]
            if abs(g) <= 1e-12:
                continue
            dn, u_ = xv[q] - lo[q], up[q] - xv[q]
            if g > 0:
                dmax, dmin = min(dmax, u_ / g), max(dmin, -dn / g)
            else:
                dmax, dmin = min(dmax, dn / -g), max(dmin, -u_ / -g)
        rf[i], rt[i] = p["rhs"][i] + dmin, p["rhs"][i] + dmax
    return of, ot, rf, rt
