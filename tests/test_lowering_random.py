"""Randomised exactness check of the for / sum_for lowering (easylp_b200/lower.py).

Random constraint bodies are drawn from the operations the trace supports (indexed variables, slices, alias rows,
parameter and loop-value coefficients, sum_for groups, + - * / with numbers and parameters, constants on both sides)
and each is built twice: by the reference's per-atom evaluation and by one symbolic evaluation.  Whenever the trace
accepts the body, the families — expanded and folded by the plain-Python oracle (oracle/lower_ref.py) — must give
the eager rows bit for bit (matrix, rhs, dir, row names).  Bodies the trace refuses are counted, not checked: falling
back is always correct.
"""
import warnings

import numpy as np
import pytest

from easylp_b200 import lower
from easylp_b200 import model as M
from fixtures import model_fold

S, T = [1, 2, 3, 4], [1, 2, 3]


class Env:
    def __init__(self, rng, api=M):
        self.api = api                   # easylp_b200.model (term lists) or oracle.dsl_ref (the reference's dense arithmetic)
        self.lp = api.easylp()
        self.x = self.lp.var("x", S, T)
        self.y = self.lp.var("y", S)
        self.z = self.lp.var("z", S, T)
        self.a = api.parameter(np.round(rng.normal(size=len(S) * len(T)), 3), S, T)
        self.w = api.parameter(np.round(rng.uniform(0.5, 2.0, len(T)), 3), T)
        self.k = api.parameter(np.round(rng.uniform(-2.0, 2.0, len(S)), 3), S)
        self.alias = api.rowSums(self.x * self.a) + 0.25


def scalar(rng, env):
    """s -> number or parameter entry (may depend on s)"""
    c = rng.integers(0, 5)
    if c == 0:
        v = float(np.round(rng.normal(), 2)) or 1.5
        return lambda s: v
    if c == 1:
        j = int(rng.integers(1, len(T) + 1))
        return lambda s: env.w[j]
    if c == 2:
        return lambda s: env.k[s]
    if c == 3:
        return lambda s: s + 0.5
    j = int(rng.integers(1, len(T) + 1))
    return lambda s: env.a[s, j] * 2


def cell(rng, env):
    """(s, t) -> the body of a sum_for"""
    c = rng.integers(0, 6)
    v = (env.x, env.z)[int(rng.integers(0, 2))]
    if c == 0:
        return lambda s, t: v[s, t]
    if c == 1:
        return lambda s, t: env.a[s, t] * v[s, t]
    if c == 2:
        return lambda s, t: v[s, t] / env.w[t] + 0.125
    if c == 3:
        j = int(rng.integers(1, len(T) + 1))
        return lambda s, t: env.w[t] * v[s, j]                  # the same column in every cell
    if c == 4:
        return lambda s, t: env.a[s, t] * env.x[s, t] - env.z[s, t] * t
    return lambda s, t: (t - 0.5) * v[s, t] + env.y[s] * env.w[t]


def node(rng, env, depth):
    """s -> a one-row expression"""
    c = rng.integers(0, 11 if depth > 0 else 5)
    if c == 0:
        return lambda s: env.y[s]
    if c == 1:
        j = int(rng.integers(1, len(T) + 1))
        v = (env.x, env.z)[int(rng.integers(0, 2))]
        return lambda s: v[s, j]
    if c == 2:
        f = cell(rng, env)
        return lambda s: env.api.sum_for(lambda t: f(s, t), t=T)
    if c == 3:
        v = (env.x, env.z)[int(rng.integers(0, 2))]
        return lambda s: env.api.Sum(v[s, :])
    if c == 4:
        return lambda s: env.alias[s]
    sub = node(rng, env, depth - 1)
    if c == 5:
        k = scalar(rng, env)
        return lambda s: k(s) * sub(s)
    if c == 6:
        k = scalar(rng, env)
        return lambda s: sub(s) / k(s)
    if c == 7:
        k = scalar(rng, env)
        return (lambda s: sub(s) + k(s)) if rng.random() < 0.5 else (lambda s: k(s) - sub(s))
    if c == 8:
        return lambda s: -sub(s)
    other = node(rng, env, depth - 1)
    if c == 9:
        return lambda s: sub(s) + other(s)
    return lambda s: sub(s) - other(s)


def build(seed, lowering, api=M):
    rng = np.random.default_rng(seed)
    old = M.LOWERING
    M.LOWERING = lowering
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = Env(rng, api)
            lhs = node(rng, env, 3)
            rhs = scalar(rng, env) if rng.random() < 0.6 else node(rng, env, 1)
            op = ("<=", ">=", "==")[int(rng.integers(0, 3))]
            body = (lambda s: lhs(s) <= rhs(s)) if op == "<=" else (lambda s: lhs(s) >= rhs(s)) if op == ">=" else \
                (lambda s: lhs(s) == rhs(s))
            env.lp.con(c=api.for_(body, s=S))
            return env.lp
    finally:
        M.LOWERING = old


SEEDS = list(range(400))


def test_random_bodies_lower_exactly_or_fall_back():
    lowered = failed = 0
    for seed in SEEDS:
        try:
            e = build(seed, False)
        except M.EasyLpError:
            failed += 1                      # e.g. a division by a parameter that is zero: the same error both ways
            with pytest.raises(M.EasyLpError):
                build(seed, True)
            continue
        l = build(seed, True)
        assert e.constraint.dir == l.constraint.dir, seed
        assert e.constraint.rhs.tobytes() == l.constraint.rhs.tobytes(), seed
        assert e.constraint.rownames == l.constraint.rownames, seed
        if not any(isinstance(b, lower.LoweredCon) for b in l._blocks):
            continue
        lowered += 1
        rp0, ci0, v0, _ = model_fold(e)
        rp, ci, v, _ = model_fold(l)
        assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0), seed
        assert v.tobytes() == v0.tobytes(), seed
    assert lowered >= len(SEEDS) // 3, (lowered, failed)


def test_random_bodies_against_the_dense_reference_arithmetic():
    """the same random bodies evaluated by the dense restatement of the reference's R arithmetic (oracle/dsl_ref.py):
    its `constraint$mat` must be what the lowered families fold to — the trace is pinned to the reference's matrix
    algebra, not only to this repo's own term-list evaluation"""
    from oracle import dsl_ref
    checked = 0
    for seed in range(120):
        try:
            d = build(seed, False, api=dsl_ref).canonical()
        except dsl_ref.RError:
            continue
        l = build(seed, True)
        rp, ci, v, _ = model_fold(l)
        assert np.array_equal(rp, d["row_ptr"]) and np.array_equal(ci, d["col_idx"]), seed
        assert v.tobytes() == d["vals"].tobytes(), seed
        assert l.constraint.rhs.tobytes() == d["rhs"].tobytes() and l.constraint.dir == d["dir"], seed
        assert l.constraint.rownames == d["rownames"], seed
        checked += any(isinstance(b, lower.LoweredCon) for b in l._blocks)
    assert checked >= 60


@pytest.mark.gpu
def test_random_bodies_on_the_device():
    """the same random bodies through elp_model_assemble: lowered families expanded and folded on the device against
    the eager term lists folded on the device"""
    checked = 0
    for seed in range(60):
        e, l = build(seed, False), build(seed, True)
        if not any(isinstance(b, lower.LoweredCon) for b in l._blocks):
            continue
        a, b = e._csr(), l._csr()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), seed
        assert np.asarray(a[2]).tobytes() == np.asarray(b[2]).tobytes(), seed
        checked += 1
    assert checked >= 30


# ---- two loop variables: row order (first index outermost), row names, coefficient tables over both -----------
def node2(rng, env, depth):
    """(s, t) -> a one-row expression"""
    c = rng.integers(0, 9 if depth > 0 else 5)
    v = (env.x, env.z)[int(rng.integers(0, 2))]
    if c == 0:
        return lambda s, t: v[s, t]
    if c == 1:
        return lambda s, t: env.y[s]
    if c == 2:
        return lambda s, t: M.sum_for(lambda u: env.a[u, t] * v[u, t], u=S)
    if c == 3:
        return lambda s, t: env.a[s, t] * v[s, t] / env.w[t]
    if c == 4:
        j = int(rng.integers(1, len(T) + 1))
        return lambda s, t: v[s, j] * (t + s / 2)
    sub = node2(rng, env, depth - 1)
    if c == 5:
        return lambda s, t: env.k[s] * sub(s, t)
    if c == 6:
        return lambda s, t: sub(s, t) - env.w[t]
    other = node2(rng, env, depth - 1)
    if c == 7:
        return lambda s, t: sub(s, t) + other(s, t)
    return lambda s, t: sub(s, t) - other(s, t)


def build2(seed, lowering, nested):
    rng = np.random.default_rng(10_000 + seed)
    old = M.LOWERING
    M.LOWERING = lowering
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = Env(rng)
            lhs = node2(rng, env, 2)
            body = lambda s, t: lhs(s, t) <= env.a[s, t] + s      # noqa: E731
            if nested:
                env.lp.con(c=M.for_(lambda s: M.for_(lambda t: body(s, t), t=T), s=S))
            else:
                env.lp.con(M.for_(body, s=S, t=T))                 # unnamed: rows "[s=1,t=1]", ...
            return env.lp
    finally:
        M.LOWERING = old


@pytest.mark.parametrize("nested", [False, True])
def test_random_bodies_over_two_indices(nested):
    lowered = 0
    for seed in range(150):
        e, l = build2(seed, False, nested), build2(seed, True, nested)
        assert e.constraint.rownames == l.constraint.rownames, seed
        assert e.constraint.names == l.constraint.names, seed
        assert e.constraint.rhs.tobytes() == l.constraint.rhs.tobytes(), seed
        if not any(isinstance(b, lower.LoweredCon) for b in l._blocks):
            continue
        lowered += 1
        rp0, ci0, v0, _ = model_fold(e)
        rp, ci, v, _ = model_fold(l)
        assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and v.tobytes() == v0.tobytes(), seed
    assert lowered >= 75
