"""Regenerates tests/golden/configs.json — run from the repo root:  python tests/golden/make_configs.py

Independent-solver (HiGHS via scipy; a stand-in for lp_solve, which is not in this image) objectives for the BASELINE
configs a CPU solver can finish: C2 (transport 300 x 300) and two small multi-commodity instances of the C5 family; the
full-size C5 (50 commodities, 5 M variables) is pinned by a combinatorial argument instead (mcnf_by_shortest_paths).
C4 carries its own planted optimum (oracle/gen.py) and needs no CPU solve."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import gen  # noqa: E402


def highs(p):
    from scipy.optimize import linprog
    from scipy.sparse import csr_matrix, vstack
    A = csr_matrix((p["vals"], p["col_idx"], p["row_ptr"]), shape=(p["m"], p["n"]))
    s = p["sense"]
    le, ge, eq = s == 0, s == 1, s == 2
    A_ub = vstack([A[le], -A[ge]]) if (le.any() or ge.any()) else None
    b_ub = np.r_[p["rhs"][le], -p["rhs"][ge]] if A_ub is not None else None
    bounds = [(None if not np.isfinite(l) else l, None if not np.isfinite(u) else u) for l, u in zip(p["lb"], p["ub"])]
    r = linprog(p["c"], A_ub=A_ub, b_ub=b_ub, A_eq=A[eq] if eq.any() else None, b_eq=p["rhs"][eq] if eq.any() else None,
                bounds=bounds, method="highs")
    assert r.status == 0
    return float(r.fun)


def mcnf_by_shortest_paths(p):
    """Exact optimum of a multi-commodity instance WITHOUT an LP solver: drop the capacity rows, the LP splits into one
    shortest-path problem per commodity (Dijkstra); if the resulting arc loads respect every capacity — asserted — that
    flow is feasible for the full LP and optimal for a relaxation of it, hence optimal.  True for the full-size C5
    (max load / capacity = 0.48), so config 5 has an objective that no first-order method produced."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra
    nodes, narcs, K = p["nodes"], p["narcs"], p["K"]
    t, h, cost, cap = p["tails"], p["heads"], p["cost"], p["cap"]
    order = np.lexsort((cost, h, t))                      # cheapest of parallel arcs first
    key = t[order] * nodes + h[order]
    first = np.ones(order.size, bool)
    first[1:] = key[1:] != key[:-1]
    sel = order[first]
    G = csr_matrix((cost[sel], (t[sel], h[sel])), shape=(nodes, nodes))
    arc_of = {(int(a), int(b)): int(i) for a, b, i in zip(t[sel], h[sel], sel)}
    dist, pred = dijkstra(G, directed=True, indices=p["src"], return_predecessors=True)
    load = np.zeros(narcs)
    obj = 0.0
    for k in range(K):
        v, s = int(p["dst"][k]), int(p["src"][k])
        obj += p["dem"][k] * dist[k, v]
        while v != s:
            u = int(pred[k, v])
            load[arc_of[(u, v)]] += p["dem"][k]
            v = u
    assert (load <= cap).all(), "capacities bind: the shortest-path decomposition is not the optimum of this instance"
    return float(obj)


def main():
    out = {
        "c5_mcnf_K50_seed0": {"objective": mcnf_by_shortest_paths(gen.mcnf(K=50)).hex(),
                              "how": "per-commodity Dijkstra; capacities verified slack (see mcnf_by_shortest_paths)"},
        "c2_transport_300x300_seed0": {"objective": highs(gen.transport(300, 300, seed=0)).hex()},
        "c5_small_K3_12x10_extra40_seed1": {"objective": highs(gen.mcnf(K=3, gw=12, gh=10, extra_arcs=40, seed=1)).hex()},
        "c5_small_K5_20x15_extra100_seed2": {"objective": highs(gen.mcnf(K=5, gw=20, gh=15, extra_arcs=100, seed=2)).hex()},
    }
    with open(os.path.join(HERE, "configs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: float.fromhex(v["objective"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
