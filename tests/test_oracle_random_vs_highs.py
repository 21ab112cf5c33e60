"""The C oracles (oracle/simplex_ref.c, oracle/pdlp_ref.c) against an independent solver on random small LPs.

HiGHS 1.12 (scipy.optimize.linprog) is the labelled stand-in for lp_solve, which is not in the image: status must agree
(optimal / infeasible / unbounded, HiGHS asked for feasibility first and for the optimum second) and optimal objectives
within 1e-7 relative.  The oracles are what the CUDA kernels are compared with, so this pins the checker itself on
instances the reference's tests do not contain: mixed senses, free and boxed variables, degenerate and infeasible rows."""
import numpy as np
import pytest
from scipy.optimize import linprog

from oracle import cbind


def random_lp(seed):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(1, 9)), int(rng.integers(1, 10))
    A = np.round(rng.normal(size=(m, n)) * 2) / 2 * (rng.random((m, n)) < 0.7)
    b = np.round(rng.normal(size=m) * 4) / 2
    sense = rng.integers(0, 3, size=m).astype(np.int8)
    c = np.round(rng.normal(size=n) * 2) / 2
    lb = np.where(rng.random(n) < 0.4, -np.inf, np.round(rng.normal(size=n)))
    ub = np.where(rng.random(n) < 0.4, np.inf, np.where(np.isfinite(lb), lb, 0.0) + np.round(rng.random(n) * 4))
    return A, b, sense, c, lb, ub


def highs(A, b, sense, c, lb, ub):
    """(lp_solve status code, objective) decided by HiGHS in two steps: feasibility with a zero objective first — with the
    objective in place HiGHS' presolve answers "infeasible" for some feasible unbounded LPs (seed 298) — then the LP."""
    le, ge, eq = sense == 0, sense == 1, sense == 2
    kw = dict(A_ub=np.vstack([A[le], -A[ge]]) if (le | ge).any() else None,
              b_ub=np.concatenate([b[le], -b[ge]]) if (le | ge).any() else None,
              A_eq=A[eq] if eq.any() else None, b_eq=b[eq] if eq.any() else None,
              bounds=list(zip(np.where(np.isfinite(lb), lb, None), np.where(np.isfinite(ub), ub, None))), method="highs")
    feas = linprog(np.zeros_like(c), **kw)
    if feas.status == 2:
        return 2, None
    if feas.status != 0:
        return None, None
    r = linprog(c, **kw)
    if r.status == 0:
        return 0, r.fun
    return (3, None) if r.status in (2, 3, 4) else (None, None)      # feasible and not optimal: unbounded


def test_simplex_oracle_matches_highs_on_random_lps():
    seen = {0: 0, 2: 0, 3: 0}
    for seed in range(400):
        A, b, sense, c, lb, ub = random_lp(seed)
        hs, hobj = highs(A, b, sense, c, lb, ub)
        if hs is None:
            continue
        st, obj, x, _ = cbind.simplex_batch(A[None], b[None], c[None], lb[None], ub[None], sense[None])
        assert int(st[0]) == hs, (seed, int(st[0]), hs)
        seen[int(st[0])] += 1
        if hs == 0:
            assert abs(obj[0] - hobj) <= 1e-7 * max(1.0, abs(hobj)), (seed, obj[0], hobj)
            ax = A @ x[0]
            viol = np.where(sense == 0, ax - b, np.where(sense == 1, b - ax, np.abs(ax - b)))
            assert viol.max(initial=0.0) <= 1e-7 and np.all(x[0] >= lb - 1e-9) and np.all(x[0] <= ub + 1e-9), seed
    assert min(seen.values()) >= 20, seen


def test_pdlp_oracle_matches_highs_on_random_feasible_lps():
    """the first-order restatement on bounded feasible instances (it needs no basis, so it is checked where it is used:
    problems with an optimum)"""
    checked = 0
    for seed in range(400, 520):
        A, b, sense, c, lb, ub = random_lp(seed)
        lb = np.where(np.isfinite(lb), lb, -3.0)
        ub = np.where(np.isfinite(ub), ub, np.maximum(lb, 0.0) + 5.0)
        hs, hobj = highs(A, b, sense, c, lb, ub)
        if hs != 0:
            continue
        m, n = A.shape
        nz = A != 0
        p = dict(m=m, n=n, row_ptr=np.r_[0, np.cumsum(nz.sum(1))].astype(np.int32), col_idx=np.nonzero(nz)[1].astype(np.int32),
                 vals=A[nz], sense=sense, rhs=b, c=c, lb=lb, ub=ub, maximize=False)
        st, out, x, y = cbind.pdlp(p, max_iter=200000)
        assert st == 0, (seed, st)
        assert abs(out[0] - hobj) <= 2e-6 * max(1.0, abs(hobj)), (seed, out[0], hobj)
        checked += 1
    assert checked >= 30
