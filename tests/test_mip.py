"""Models with integer / binary variables: `solve(prob)` after `set.type(...)` is lp_solve's branch and bound in the
reference (/root/reference/R/class.R:264-276).  CPU: the oracle B&B (oracle/mip_ref.py over the C simplex oracle) against
the values the reference's own tests pin and against HiGHS' branch and cut on random MILPs.  GPU: elp_solve_mip
(csrc/mip.cu: frontiers of the tree through the batched simplex kernel) against the oracle — status exact, objective
within 1e-9 relative."""
import json
import os

import numpy as np
import pytest

import models
from easylp_b200 import _lib as L
from easylp_b200 import model as M
from oracle import dsl_ref, mip_ref

HERE = os.path.dirname(os.path.abspath(__file__))


def _random_milp(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(3, 11)), int(rng.integers(2, 9))
    A = np.round(rng.uniform(-3, 6, size=(m, n))) * (rng.random((m, n)) < 0.8)
    x0 = rng.integers(0, 5, size=n).astype(float)                   # a feasible integer point
    sense = rng.choice([0, 0, 1, 2], size=m).astype(np.int8)
    slack = rng.integers(0, 6, size=m)
    rhs = A @ x0 + np.where(sense == 0, slack, np.where(sense == 1, -slack, 0))
    c = np.round(rng.uniform(-5, 5, size=n))
    lb = np.zeros(n)
    ub = np.where(rng.random(n) < 0.8, rng.integers(4, 12, size=n), 30).astype(float)
    is_int = (rng.random(n) < 0.7).astype(np.uint8)
    is_int[0] = 1
    rp = np.zeros(m + 1, np.int32)
    ci, v = [], []
    for i in range(m):
        nz = np.nonzero(A[i])[0]
        ci += nz.tolist(); v += A[i, nz].tolist()
        rp[i + 1] = len(ci)
    return dict(m=m, n=n, row_ptr=rp, col_idx=np.array(ci, np.int32), vals=np.array(v, float), sense=sense, rhs=rhs, c=c,
                lb=lb, ub=ub, is_integer=is_int, maximize=bool(rng.random() < 0.5))


def _highs(p):
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import csr_matrix
    A = csr_matrix((p["vals"], p["col_idx"], p["row_ptr"]), shape=(p["m"], p["n"]))
    sign = -1.0 if p["maximize"] else 1.0
    lo = np.where(p["sense"] == 0, -np.inf, p["rhs"])
    hi = np.where(p["sense"] == 1, np.inf, p["rhs"])
    r = milp(sign * p["c"], constraints=LinearConstraint(A, lo, hi), integrality=p["is_integer"].astype(int),
             bounds=Bounds(p["lb"], p["ub"]))
    return {0: 0, 2: 2, 3: 3}.get(r.status, -1), (sign * r.fun if r.status == 0 else None)


def _oracle(p, **kw):
    return mip_ref.solve(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
                         p["is_integer"], maximize=p["maximize"], **kw)


def test_oracle_reproduces_the_values_the_reference_tests_pin():
    # test-investments.R:45-46
    can = models.ALL["investments_assembly"](dsl_ref).canonical()
    st, obj, x, _ = _oracle(can)
    assert st == 0 and obj == 469.0 and x.tolist() == [0, 0, 1, 1, 1, 0]
    # test-cyingair.R:27-30: x = (0, 2, 3, 49), quin = (0, 1, 1, 1); variables are (quin, x) in declaration order
    can = models.ALL["cyingair"](dsl_ref).canonical()
    st, obj, x, _ = _oracle(can)
    assert st == 0 and np.allclose(x, [0, 1, 1, 1, 0, 2, 3, 49], atol=1e-7)
    assert abs(obj - (5.8 * 0 + 4.2 * 2 + 3 * 3 + 2.3 * 49)) <= 1e-9


def test_oracle_against_the_committed_milp_goldens():
    from fixtures import load_golden
    gold = {k: v for k, v in load_golden().items() if "highs_milp" in v}
    assert set(gold) >= {"cyingair", "investments_assembly"}
    for name, g in gold.items():
        st, obj, x, _ = _oracle(g)
        assert st == g["highs_milp"]["status"]
        assert abs(obj - g["highs_milp"]["objective"]) <= 1e-9 * max(1.0, abs(obj))


def test_oracle_against_highs_on_random_milps():
    seen = {0: 0, 2: 0}
    for seed in range(120):
        p = _random_milp(seed)
        st, obj, x, nodes = _oracle(p)
        hs, ho = _highs(p)
        assert st == hs, (seed, st, hs)
        seen[st] = seen.get(st, 0) + 1
        if st == 0:
            assert abs(obj - ho) <= 1e-7 * max(1.0, abs(ho)), (seed, obj, ho)
            assert np.all(np.abs(x[p["is_integer"] == 1] - np.round(x[p["is_integer"] == 1])) <= 1e-7)
    assert seen[0] >= 80


def test_oracle_infeasible_and_unbounded_milp():
    # x + y <= 1.5, x + y >= 1.2, both integer in [0, 5]: the relaxation is feasible, the MILP is not
    p = dict(m=2, n=2, row_ptr=np.array([0, 2, 4], np.int32), col_idx=np.array([0, 1, 0, 1], np.int32), vals=np.ones(4),
             sense=np.array([0, 1], np.int8), rhs=np.array([1.5, 1.2]), c=np.array([1.0, 1.0]), lb=np.zeros(2),
             ub=np.full(2, 5.0), is_integer=np.array([1, 1], np.uint8), maximize=False)
    assert _oracle(p)[0] == 2
    q = dict(p, ub=np.full(2, np.inf), sense=np.array([1, 1], np.int8), maximize=True)
    assert _oracle(q)[0] == 3


# ---- GPU ------------------------------------------------------------------------------------------------
def _gpu(p, **kw):
    return L.solve_mip(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
                       p["is_integer"], maximize=p["maximize"], options=L.default_options(**kw) if kw else None)


@pytest.mark.gpu
def test_gpu_branch_and_bound_matches_the_reference_pinned_solutions():
    lp = models.ALL["investments_assembly"](M)
    lp.solve()
    assert lp.status == "optimal" and lp.objective_value == 469.0                 # test-investments.R:45
    assert np.asarray(lp.solution["x"]).ravel().tolist() == [0, 0, 1, 1, 1, 0]    # test-investments.R:46
    lp = models.ALL["cyingair"](M)
    lp.solve()
    assert lp.status == "optimal"
    assert np.allclose(np.asarray(lp.solution["x"]).ravel(), [0, 2, 3, 49], atol=1e-7)      # test-cyingair.R:28
    assert np.allclose(np.asarray(lp.solution["quin"]).ravel(), [0, 1, 1, 1], atol=1e-7)    # test-cyingair.R:29


@pytest.mark.gpu
def test_gpu_branch_and_bound_against_the_oracle_on_random_milps():
    for seed in range(150):
        p = _random_milp(seed)
        st, obj, x, nodes = _oracle(p)
        r = _gpu(p)
        assert r.status == st, (seed, r.status, st)
        if st == 0:
            assert abs(r.objval - obj) <= 1e-9 * max(1.0, abs(obj)), (seed, r.objval, obj)
            ii = p["is_integer"] == 1
            assert np.all(np.abs(r.x[ii] - np.round(r.x[ii])) <= 1e-7)
            assert np.all(r.x >= p["lb"] - 1e-9) and np.all(r.x <= p["ub"] + 1e-9)
            assert r.stats.restarts >= 1           # nodes solved


@pytest.mark.gpu
def test_gpu_milp_statuses_and_limits():
    p = dict(m=2, n=2, row_ptr=np.array([0, 2, 4], np.int32), col_idx=np.array([0, 1, 0, 1], np.int32), vals=np.ones(4),
             sense=np.array([0, 1], np.int8), rhs=np.array([1.5, 1.2]), c=np.array([1.0, 1.0]), lb=np.zeros(2),
             ub=np.full(2, 5.0), is_integer=np.array([1, 1], np.uint8), maximize=False)
    assert _gpu(p).status == L.STATUS_INFEASIBLE
    q = dict(p, ub=np.full(2, np.inf), sense=np.array([1, 1], np.int8), maximize=True)
    r = _gpu(q)
    assert r.status == L.STATUS_UNBOUNDED and r.objval == np.inf
    # a node limit of 1 stops after the root: no incumbent yet -> timeout (7), lp_solve's code for a stopped search
    hard = _random_milp(7)
    full = _gpu(hard)
    if full.stats.restarts > 1:
        cut = _gpu(hard, max_iter=1)
        assert cut.status in (L.STATUS_TIMEOUT, 1) and cut.stats.limit_reached == 1
