"""Seeded synthetic generators for the BASELINE.json configs.

TEST INFRASTRUCTURE + bench input generation.  Pure numpy.  These produce the
*inputs* (CSR + vectors) handed identically to the CUDA path and to the CPU
oracle; they are not part of the solve path.

Row sense encoding used throughout the repo (matches include/easylp_abi.h):
    0 : "<="    1 : ">="    2 : "=="
"""
from __future__ import annotations

import numpy as np

LE, GE, EQ = 0, 1, 2


def readme_lp():
    """G1 — /root/reference/README.md:16-24.  max x+y; x+2y<=3; y>=3x-2 -> -3x+y>=-2."""
    row_ptr = np.array([0, 2, 4], dtype=np.int32)
    col_idx = np.array([0, 1, 0, 1], dtype=np.int32)
    vals = np.array([1.0, 2.0, -3.0, 1.0])
    sense = np.array([LE, GE], dtype=np.int8)
    rhs = np.array([3.0, -2.0])
    c = np.array([1.0, 1.0])
    lb = np.full(2, -np.inf)
    ub = np.full(2, np.inf)
    return dict(m=2, n=2, row_ptr=row_ptr, col_idx=col_idx, vals=vals, sense=sense,
                rhs=rhs, c=c, lb=lb, ub=ub, maximize=True)


def transport(S=300, T=300, seed=0):
    """C2 — transportation problem, rows in the `$con` order of SURVEY §8(d):
    S supply rows  sum_t x[s,t] <= supply[s], then T demand rows sum_s x[s,t] >= demand[t].
    Column ids are column-major over (s,t) like reference R/class.R:112-113:
    col(s,t) = s + S*t."""
    rng = np.random.default_rng(seed)
    cost = rng.uniform(10, 90, size=(S, T))
    demand = rng.uniform(20, 80, size=T)
    supply = rng.uniform(50, 150, size=S)
    supply *= 1.1 * demand.sum() / supply.sum()
    n = S * T
    m = S + T
    row_ptr = np.zeros(m + 1, dtype=np.int32)
    row_ptr[1:S + 1] = T
    row_ptr[S + 1:] = S
    row_ptr = np.cumsum(row_ptr, dtype=np.int64).astype(np.int32)
    s_idx = np.arange(S)
    t_idx = np.arange(T)
    sup_cols = (s_idx[:, None] + S * t_idx[None, :]).astype(np.int32)      # row s: cols over t (ascending)
    dem_cols = (s_idx[None, :] + S * t_idx[:, None]).astype(np.int32)      # row t: cols over s
    col_idx = np.concatenate([sup_cols.ravel(), dem_cols.ravel()])
    vals = np.ones(col_idx.size)
    sense = np.concatenate([np.full(S, LE), np.full(T, GE)]).astype(np.int8)
    rhs = np.concatenate([supply, demand])
    c = cost.ravel(order="F").copy()    # column-major like the variable ids
    return dict(m=m, n=n, row_ptr=row_ptr, col_idx=col_idx, vals=vals, sense=sense, rhs=rhs,
                c=c, lb=np.zeros(n), ub=np.full(n, np.inf), maximize=False,
                cost=cost, supply=supply, demand=demand)


def dense_batch(B=200_000, m=20, n=30, seed=0):
    """C3 — batch of random feasible bounded dense LPs (min c'x, A x <= b, 0<=x<=10).
    Layout: A[B][m][n] row-major, b[B][m], c[B][n]."""
    rng = np.random.default_rng(seed)
    A = rng.uniform(0, 1, size=(B, m, n))
    x0 = rng.uniform(0, 1, size=(B, n))
    b = np.einsum("bij,bj->bi", A, x0) + rng.uniform(0.1, 1.0, size=(B, m))
    c = -rng.uniform(0, 1, size=(B, n))
    lb = np.zeros((B, n))
    ub = np.full((B, n), 10.0)
    sense = np.zeros((B, m), dtype=np.int8)
    return dict(B=B, m=m, n=n, A=A, b=b, c=c, lb=lb, ub=ub, sense=sense)


def sparse_planted(m=2_000_000, n=None, seed=0, ub_val=1e3, frac_basic=0.3, frac_active=0.6):
    """C4 — synthetic sparse LP with a planted primal-dual optimal pair.

    Row lengths ~ U{7..13}; columns uniform (duplicates within a row are nudged apart so each
    row has distinct, ascending columns); values ~ U(-1,1) bounded away from 0.
    Senses: 50% <=, 25% >=, 25% ==.  min c'x, 0 <= x <= ub_val.
    Planted pair (x0, y0): x0_j > 0 on a `frac_basic` subset (strictly inside the box), rows are
    active (slack 0) with prob `frac_active` and carry a nonzero multiplier of the right sign,
    c = A'y0 + r with r_j >= 0 complementary to x0 (r_j = 0 where x0_j > 0).  => c'x0 is the optimum.
    Sign convention: y_i >= 0 for >= rows, y_i <= 0 for <= rows (Lagrangian c'x - y'(Ax - b))."""
    if n is None:
        n = 2 * m
    rng = np.random.default_rng(seed)
    lens = rng.integers(7, 14, size=m).astype(np.int64)
    row_ptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=row_ptr[1:])
    nnz = int(row_ptr[-1])
    rows = np.repeat(np.arange(m, dtype=np.int64), lens)
    cols = rng.integers(0, n, size=nnz).astype(np.int64)
    # sort within rows, then de-duplicate by nudging (keeps ascending & distinct; wraps rarely)
    key = rows * n + cols
    key.sort()
    dup = np.zeros(nnz, dtype=bool)
    dup[1:] = key[1:] == key[:-1]
    while dup.any():
        key[dup] += 1
        # keep inside the row
        over = (key // n) != rows
        key[over] = rows[over] * n + rng.integers(0, n, size=int(over.sum()))
        order = np.argsort(key, kind="stable")
        key = key[order]
        dup[:] = False
        dup[1:] = key[1:] == key[:-1]
    cols = (key - rows * n).astype(np.int32)
    vals = rng.uniform(0.05, 1.0, size=nnz) * rng.choice([-1.0, 1.0], size=nnz)

    sense = rng.choice(np.array([LE, LE, GE, EQ], dtype=np.int8), size=m)
    x0 = np.zeros(n)
    basic = rng.random(n) < frac_basic
    x0[basic] = rng.uniform(0.5, 5.0, size=int(basic.sum()))
    ax0 = _csr_matvec(row_ptr, cols, vals, x0, m)
    active = (rng.random(m) < frac_active) | (sense == EQ)
    slack = np.where(active, 0.0, rng.uniform(0.5, 2.0, size=m))
    rhs = np.where(sense == LE, ax0 + slack, np.where(sense == GE, ax0 - slack, ax0))
    ymag = rng.uniform(0.1, 1.0, size=m)
    y0 = np.where(sense == LE, -ymag, np.where(sense == GE, ymag, ymag * rng.choice([-1.0, 1.0], size=m)))
    y0[~active] = 0.0
    aty0 = _csr_rmatvec(row_ptr, cols, vals, y0, n)
    r = np.where(basic, 0.0, rng.uniform(0.1, 1.0, size=n))
    c = aty0 + r
    return dict(m=m, n=n, row_ptr=row_ptr.astype(np.int32), col_idx=cols, vals=vals, sense=sense, rhs=rhs,
                c=c, lb=np.zeros(n), ub=np.full(n, ub_val), maximize=False,
                x_opt=x0, y_opt=y0, obj_opt=float(c @ x0))


def _csr_matvec(row_ptr, cols, vals, x, m):
    prod = vals * x[cols]
    out = np.add.reduceat(prod, row_ptr[:-1].astype(np.int64)) if prod.size else np.zeros(m)
    lens = np.diff(row_ptr)
    out = np.where(lens > 0, out, 0.0)
    return out


def _csr_rmatvec(row_ptr, cols, vals, y, n):
    lens = np.diff(row_ptr)
    rows = np.repeat(np.arange(len(lens)), lens)
    return np.bincount(cols, weights=vals * y[rows], minlength=n)


def mcnf(K=50, gw=100, gh=200, extra_arcs=20_600, seed=0):
    """C5 — multi-commodity network flow on a gw x gh grid.
    Variables x[k, a] (col = k*narcs + a).  Rows: K*nodes conservation (==) then narcs capacity (<=).
    Feasibility is ensured by routing every commodity along a path first and sizing capacities
    above the resulting load."""
    rng = np.random.default_rng(seed)
    nodes = gw * gh
    nid = np.arange(nodes).reshape(gh, gw)
    tails, heads = [], []
    for (a, b) in ((nid[:, :-1], nid[:, 1:]), (nid[:, 1:], nid[:, :-1]),
                   (nid[:-1, :], nid[1:, :]), (nid[1:, :], nid[:-1, :])):
        tails.append(a.ravel()); heads.append(b.ravel())
    tails = np.concatenate(tails); heads = np.concatenate(heads)
    et = rng.integers(0, nodes, size=extra_arcs)
    eh = (et + rng.integers(1, nodes, size=extra_arcs)) % nodes
    tails = np.concatenate([tails, et]).astype(np.int64)
    heads = np.concatenate([heads, eh]).astype(np.int64)
    narcs = tails.size
    cost = rng.uniform(1, 10, size=narcs)
    src = rng.integers(0, nodes, size=K)
    dst = (src + rng.integers(1, nodes, size=K)) % nodes
    dem = rng.uniform(1, 10, size=K)
    # grid arc lookup for routing (row-then-column Manhattan path)
    W = gw
    nh = gh * (gw - 1)          # right arcs count; left arcs next; then down, up
    nv = (gh - 1) * gw
    def right(r, ccol): return r * (gw - 1) + ccol
    def left(r, ccol): return nh + r * (gw - 1) + (ccol - 1)
    def down(r, ccol): return 2 * nh + r * gw + ccol
    def up(r, ccol): return 2 * nh + nv + (r - 1) * gw + ccol
    load = np.zeros(narcs)
    for k in range(K):
        r0, c0 = divmod(int(src[k]), W); r1, c1 = divmod(int(dst[k]), W)
        cc = c0
        while cc < c1: load[right(r0, cc)] += dem[k]; cc += 1
        while cc > c1: load[left(r0, cc)] += dem[k]; cc -= 1
        rr = r0
        while rr < r1: load[down(rr, c1)] += dem[k]; rr += 1
        while rr > r1: load[up(rr, c1)] += dem[k]; rr -= 1
    cap = np.maximum(rng.uniform(20, 60, size=narcs), 1.25 * load)
    n = K * narcs
    m = K * nodes + narcs
    # conservation rows: for node v, commodity k: sum_out x - sum_in x = supply
    # build COO then CSR (row-major, ascending col)
    a_idx = np.arange(narcs, dtype=np.int64)
    rows_l, cols_l, vals_l = [], [], []
    for k in range(K):
        rows_l.append(k * nodes + tails); cols_l.append(k * narcs + a_idx); vals_l.append(np.ones(narcs))
        rows_l.append(k * nodes + heads); cols_l.append(k * narcs + a_idx); vals_l.append(-np.ones(narcs))
    kk = np.repeat(np.arange(K, dtype=np.int64), narcs)
    rows_l.append(K * nodes + np.tile(a_idx, K)); cols_l.append(kk * narcs + np.tile(a_idx, K)); vals_l.append(np.ones(n))
    rows = np.concatenate(rows_l); cols = np.concatenate(cols_l); vals = np.concatenate(vals_l)
    order = np.lexsort((cols, rows))
    rows = rows[order]; cols = cols[order]; vals = vals[order]
    row_ptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=m), out=row_ptr[1:])
    rhs = np.zeros(m)
    for k in range(K):
        rhs[k * nodes + src[k]] += dem[k]
        rhs[k * nodes + dst[k]] -= dem[k]
    rhs[K * nodes:] = cap
    sense = np.concatenate([np.full(K * nodes, EQ), np.full(narcs, LE)]).astype(np.int8)
    c = np.tile(cost, K)
    return dict(m=m, n=n, row_ptr=row_ptr.astype(np.int32), col_idx=cols.astype(np.int32), vals=vals,
                sense=sense, rhs=rhs, c=c, lb=np.zeros(n), ub=np.full(n, np.inf), maximize=False,
                K=K, nodes=nodes, narcs=narcs, tails=tails, heads=heads, cost=cost, cap=cap, src=src, dst=dst, dem=dem)


def term_stream(p, dup_frac=0.05, seed=0):
    """An emission-order term list whose canonical CSR is exactly p's matrix: every entry once, in a seeded random order,
    with a fraction of the entries split into two halves (v/2 + v/2 folds back to v exactly) so that the ordered
    duplicate fold of elp_assemble_csr has work to do.  Returns (row, col, val) int32/int32/float64."""
    m = p["m"]
    nnz = int(p["row_ptr"][m])
    rows = np.repeat(np.arange(m, dtype=np.int32), np.diff(p["row_ptr"]))
    rng = np.random.default_rng(seed)
    dup = rng.random(nnz) < dup_frac
    r = np.concatenate([rows, rows[dup]])
    c = np.concatenate([p["col_idx"], p["col_idx"][dup]])
    v = np.concatenate([np.where(dup, p["vals"] * 0.5, p["vals"]), p["vals"][dup] * 0.5])
    perm = rng.permutation(r.size)
    return r[perm], c[perm].astype(np.int32), v[perm]
