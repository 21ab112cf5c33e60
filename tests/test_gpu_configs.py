"""GPU: the five BASELINE.json configs at FULL size through the product path, checked by size-independent properties
(known planted optimum, independent-solver objective where a CPU solver can finish, residual / feasibility /
duality checks recomputed on the host) — north_star: status exact, objective <= 1e-6 relative, residuals <= 1e-6."""
import json
import os
import warnings

import numpy as np
import pytest

from easylp_b200 import _lib as L
from easylp_b200 import model as M
from oracle import cbind, gen

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "configs.json")) as f:
    CONFIGS = {k: float.fromhex(v["objective"]) for k, v in json.load(f).items()}


def _pdlp(p, **kw):
    return L.solve_lp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
                      maximize=p.get("maximize", False), options=L.default_options(method=L.METHOD_PDLP, **kw))


def _host_residuals(p, x):
    """relative primal residual recomputed on the host from the returned x"""
    ax = gen._csr_matvec(p["row_ptr"].astype(np.int64), p["col_idx"], p["vals"], x, p["m"])
    lc = np.where(p["sense"] == 0, -np.inf, p["rhs"])
    uc = np.where(p["sense"] == 1, np.inf, p["rhs"])
    viol = ax - np.clip(ax, lc, uc)
    nb = np.linalg.norm(p["rhs"])
    return np.linalg.norm(viol) / (1 + nb), np.maximum(p["lb"] - x, 0).max(), np.maximum(x - p["ub"], 0).max()


def test_c2_transport_300x300_through_the_dsl():
    """config 2: built with for + sum_for on the host DSL, assembled on the device, solved by PDLP"""
    S = T = 300
    p = gen.transport(S, T, seed=0)
    src, snk = list(range(1, S + 1)), list(range(1, T + 1))
    lp = M.easylp()
    x = lp.var("x", src, snk, lower=0)
    cost = M.parameter(p["cost"].ravel(order="F"), src, snk)
    supply, demand = M.parameter(p["supply"], src), M.parameter(p["demand"], snk)
    lp.min(M.Sum(cost * x))
    lp.con(make=M.for_(lambda s: M.sum_for(lambda t: x[s, t], t=snk) <= supply[s], s=src),
           sell=M.for_(lambda t: M.sum_for(lambda s: x[s, t], s=src) >= demand[t], t=snk))
    rp, ci, v = lp._csr()
    # device assembly reproduces the generator's canonical CSR bit for bit (90 000 vars, 600 rows, 180 000 nnz)
    assert np.array_equal(rp, p["row_ptr"]) and np.array_equal(ci, p["col_idx"]) and np.asarray(v).tobytes() == p["vals"].tobytes()
    assert lp.objective_fun.tobytes() == p["c"].tobytes()
    lp.solve(gpu_method="pdlp")
    ref = CONFIGS["c2_transport_300x300_seed0"]
    assert lp.status == "optimal"
    assert abs(lp.objective_value - ref) <= 1e-6 * abs(ref)
    st = lp.pointer
    assert st.method_used == L.METHOD_PDLP
    assert st.rel_primal_res <= 1e-6 and st.rel_dual_res <= 1e-6 and st.rel_gap <= 1e-6
    xs = lp.solution["x"].ravel(order="F")
    pres, lo, hi = _host_residuals(p, xs)
    assert pres <= 2e-6 and lo <= 1e-9


def test_c3_batch_200k():
    """config 3: 200 000 feasible dense LPs of 20 x 30"""
    d = gen.dense_batch(B=200_000, seed=0)
    status, obj, x, st = L.solve_batch(d["A"], d["b"], d["c"], d["lb"], d["ub"], d["sense"])
    assert np.all(status == 0)
    ax = np.einsum("bij,bj->bi", d["A"], x)
    assert np.all(ax <= d["b"] + 1e-7 * np.maximum(1.0, np.abs(d["b"])))
    assert np.all(x >= d["lb"] - 1e-9) and np.all(x <= d["ub"] + 1e-9)
    assert np.allclose(obj, np.einsum("bj,bj->b", d["c"], x), rtol=1e-9, atol=1e-9)
    # a spread sample against the CPU oracle: status exact, objective <= 1e-6 relative
    idx = np.arange(0, 200_000, 97)
    s0, o0, _, _ = cbind.simplex_batch(d["A"][idx], d["b"][idx], d["c"][idx], d["lb"][idx], d["ub"][idx], d["sense"][idx],
                                       nthreads=8)
    assert np.array_equal(status[idx], s0)
    assert np.all(np.abs(obj[idx] - o0) <= 1e-6 * np.maximum(1.0, np.abs(o0)))


def test_c4_sparse_2m_x_4m_planted_optimum():
    """config 4: the optimum is planted by the generator, so the objective is known without a CPU solver"""
    p = gen.sparse_planted(2_000_000, seed=0)
    r = _pdlp(p)
    assert r.status == 0
    assert abs(r.objval - p["obj_opt"]) <= 1e-6 * max(1.0, abs(p["obj_opt"]))
    assert r.stats.rel_primal_res <= 1e-6 and r.stats.rel_dual_res <= 1e-6 and r.stats.rel_gap <= 1e-6
    pres, lo, hi = _host_residuals(p, r.x)
    assert pres <= 2e-6 and lo <= 1e-9 and hi <= 1e-9
    assert abs(float(np.dot(p["c"], r.x)) - r.objval) <= 1e-9 * max(1.0, abs(r.objval))


@pytest.mark.parametrize("key,kw", [("c5_small_K3_12x10_extra40_seed1", dict(K=3, gw=12, gh=10, extra_arcs=40, seed=1)),
                                    ("c5_small_K5_20x15_extra100_seed2", dict(K=5, gw=20, gh=15, extra_arcs=100, seed=2))])
def test_c5_family_small_vs_independent_solver(key, kw):
    p = gen.mcnf(**kw)
    r = _pdlp(p)
    ref = CONFIGS[key]
    assert r.status == 0 and abs(r.objval - ref) <= 1e-6 * max(1.0, abs(ref))


def test_c5_multicommodity_50_commodities():
    """config 5: 50 commodities on 20 000 nodes / 100 000 arcs (5 M variables, 1.1 M rows)"""
    p = gen.mcnf(K=50)
    assert (p["n"], p["m"]) == (5_000_000, 1_100_000)
    r = _pdlp(p)
    assert r.status == 0
    assert r.stats.rel_primal_res <= 1e-6 and r.stats.rel_dual_res <= 1e-6 and r.stats.rel_gap <= 1e-6
    pres, lo, hi = _host_residuals(p, r.x)
    assert pres <= 2e-6 and lo <= 1e-9
    # weak duality with the returned duals: b'y (+ bound terms are zero here: lb = 0, ub = inf) <= c'x + gap
    y = r.y
    dual = float(np.dot(np.where(np.isfinite(p["rhs"]), p["rhs"], 0.0), y))
    assert dual <= r.objval + 1e-6 * (1 + abs(r.objval) + abs(dual))
    assert abs(r.objval - dual) <= 2e-6 * (1 + abs(r.objval) + abs(dual))
    # independent objective: the instance decomposes into 50 shortest-path problems whose loads respect every capacity
    # (tests/golden/make_configs.py::mcnf_by_shortest_paths) -- no first-order method produced this number
    ref = CONFIGS["c5_mcnf_K50_seed0"]
    assert abs(r.objval - ref) <= 1e-6 * abs(ref)
    # dual feasibility recomputed on the host: reduced costs c - A'y may only be negative within the tolerance
    # (lb = 0, ub = +inf), and the capacity rows (<=) carry non-positive multipliers in this sign convention
    rc = p["c"] - gen._csr_rmatvec(p["row_ptr"].astype(np.int64), p["col_idx"], p["vals"], y, p["n"])
    assert np.linalg.norm(np.minimum(rc, 0.0)) <= 1e-6 * (1 + np.linalg.norm(p["c"]))
    le = p["sense"] == 0
    assert np.linalg.norm(np.maximum(y[le], 0.0)) <= 1e-6 * (1 + np.linalg.norm(y))
