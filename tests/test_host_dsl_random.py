"""Randomised check of the host DSL's vector algebra (easylp_b200.model, sparse term lists) against the dense
restatement of the reference's R arithmetic (oracle/dsl_ref.py): random multi-row expressions built from `[` with index
vectors, recycling products and quotients, sums and differences of variables, `sum`/`mean`/`cumsum`/`rowSums`/
`colSums`, and comparisons with vector right-hand sides.  Matrix, rhs, dir and row names must agree bit for bit; an
error in one must be an error in the other."""
import warnings

import numpy as np
import pytest

from easylp_b200 import model as M
from fixtures import model_fold
from oracle import dsl_ref

S, T = [1, 2, 3, 4], [1, 2, 3]


def vec(rng, api, env, depth):
    """-> (expression, nrow) ; expressions are multi-row lp_vars"""
    x, y, z, a = env
    c = int(rng.integers(0, 12 if depth > 0 else 6))
    if c == 0:
        return x, 12
    if c == 1:
        return y, 4
    if c == 2:
        j = int(rng.integers(1, 4))
        return x[:, j], 4
    if c == 3:
        i = int(rng.integers(1, 5))
        return z[i, :], 3
    if c == 4:
        rows = sorted(set(int(v) for v in rng.integers(1, 5, size=3)))
        return y[rows], len(rows)
    if c == 5:
        return api.rowSums(x * a), 4
    e, n = vec(rng, api, env, depth - 1)
    if c == 6:
        k = np.round(rng.normal(size=n), 2)
        return e * k, n
    if c == 7:
        return e / float(np.round(rng.uniform(0.5, 3.0), 2)), n
    if c == 8:
        return -e + float(np.round(rng.normal(), 2)), n
    if c == 9:
        return api.cumsum(e), n
    f, m = vec(rng, api, env, depth - 1)
    if c == 10:
        return (e + f, max(n, m)) if (n == m or 1 in (n, m)) else (e, n)
    return (e - f, max(n, m)) if (n == m or 1 in (n, m)) else (f, m)


def build(seed, api):
    rng = np.random.default_rng(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lp = api.easylp()
        x, y, z = lp.var("x", S, T), lp.var("y", S, lower=0), lp.var("z", S, T)
        a = api.parameter(np.round(rng.normal(size=12), 3), S, T)
        env = (x, y, z, a)
        e, n = vec(rng, api, env, 3)
        kind = int(rng.integers(0, 4))
        if kind == 0:
            con = e <= np.round(rng.normal(size=n), 2)
        elif kind == 1:
            con = api.Sum(e, y[1]) >= float(np.round(rng.normal(), 2))
        elif kind == 2:
            con = api.mean(e) == 1.5
        else:
            f, m = vec(rng, api, env, 1)
            con = (e >= f) if (n == m or 1 in (n, m)) else (e <= 0)
        lp.con(r=con)
        return lp


def test_random_vector_expressions_match_the_dense_arithmetic():
    checked = errors = 0
    for seed in range(300):
        try:
            d = build(seed, dsl_ref).canonical()
        except dsl_ref.RError:
            errors += 1
            with pytest.raises(M.EasyLpError):
                build(seed, M)
            continue
        lp = build(seed, M)
        rp, ci, v, m = model_fold(lp)
        assert m == d["m"], seed
        assert np.array_equal(rp, d["row_ptr"]) and np.array_equal(ci, d["col_idx"]), seed
        assert v.tobytes() == d["vals"].tobytes(), seed
        assert lp.constraint.rhs.tobytes() == d["rhs"].tobytes() and lp.constraint.dir == d["dir"], seed
        assert lp.constraint.rownames == d["rownames"], seed
        checked += 1
    assert checked >= 250, (checked, errors)
