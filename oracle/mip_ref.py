"""CPU restatement of the branch and bound of easylp_b200/csrc/mip.cu.  TEST INFRASTRUCTURE ONLY.

Stands where the reference calls `solve(prob)` on a model with `set.type(prob, columns, "integer" | "binary")`
(/root/reference/R/class.R:264-276: lp_solve's branch and bound).  Node LPs are solved by the C simplex oracle
(oracle/simplex_ref.c through cbind.simplex_csr); the tree is walked depth-first, branching on the lowest-indexed
fractional integer column (lp_solve's default NODE_FIRSTSELECT), integrality tolerance 1e-7 (lp_solve's epsint),
pruning gaps 1e-11 absolute / 1e-9 relative (lp_solve's mip_gap defaults).

Pinned by the values the reference's own tests hold: test-investments.R:45-46 (objective 469, x = 0 0 1 1 1 0) and
test-cyingair.R:27-30 (x = 0 2 3 49, quin = 0 1 1 1) — tests/test_oracle_golden.py — and by HiGHS' branch and cut on
random small MILPs (tests/test_mip.py)."""
from __future__ import annotations

import math

import numpy as np

from . import cbind

EPS_INT, GAP_ABS, GAP_REL = 1e-7, 1e-11, 1e-9


def solve(m, n, row_ptr, col_idx, vals, sense, rhs, c, lb, ub, is_integer, maximize=False, node_limit=200_000):
    """returns (status, objective, x, nodes) with lp_solve status codes: 0 optimal, 2 unfeasible, 3 unbounded, 1 node limit"""
    lb = np.array(np.broadcast_to(lb, (n,)), dtype=float)
    ub = np.array(np.broadcast_to(ub, (n,)), dtype=float)
    is_integer = np.asarray(is_integer).astype(bool)
    for j in np.nonzero(is_integer)[0]:
        if np.isfinite(lb[j]):
            lb[j] = math.ceil(lb[j] - EPS_INT)
        if np.isfinite(ub[j]):
            ub[j] = math.floor(ub[j] + EPS_INT)
    best, best_x, nodes = math.inf, None, 0
    stack = [(lb, ub)]
    while stack:
        if nodes >= node_limit:
            return (1 if best_x is not None else 7), (-best if maximize else best), best_x, nodes
        l, u = stack.pop()
        nodes += 1
        if np.any(l > u):
            continue
        st, obj, x, _, _ = cbind.simplex_csr(m, n, row_ptr, col_idx, vals, sense, rhs, c, l, u, maximize=maximize)
        if st == 2:
            continue
        if st == 3:
            return 3, (math.inf if maximize else -math.inf), None, nodes
        if st != 0:
            return 5, 0.0, None, nodes
        v = -obj if maximize else obj
        if v >= best - max(GAP_ABS, GAP_REL * abs(best)) if best_x is not None else False:
            continue
        frac = [j for j in range(n) if is_integer[j] and abs(x[j] - round(x[j])) > EPS_INT]
        if not frac:
            best, best_x = v, x.copy()
            continue
        j = frac[0]
        up_l = l.copy(); up_l[j] = math.ceil(x[j])
        dn_u = u.copy(); dn_u[j] = math.floor(x[j])
        stack.append((up_l, u))          # explored second
        stack.append((l, dn_u))          # floor branch first (lp_solve's default floor_first = BRANCH_CEILING off)
    if best_x is None:
        return 2, 0.0, None, nodes
    return 0, (-best if maximize else best), best_x, nodes
