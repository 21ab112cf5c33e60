"""Regenerates tests/golden/configs.json — run from the repo root:  python tests/golden/make_configs.py

Independent-solver (HiGHS via scipy; a stand-in for lp_solve, which is not in this image) objectives for the BASELINE
configs a CPU solver can finish: C2 (transport 300 x 300) and a small multi-commodity instance of the C5 family.
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


def main():
    out = {
        "c2_transport_300x300_seed0": {"objective": highs(gen.transport(300, 300, seed=0)).hex()},
        "c5_small_K3_12x10_extra40_seed1": {"objective": highs(gen.mcnf(K=3, gw=12, gh=10, extra_arcs=40, seed=1)).hex()},
        "c5_small_K5_20x15_extra100_seed2": {"objective": highs(gen.mcnf(K=5, gw=20, gh=15, extra_arcs=100, seed=2)).hex()},
    }
    with open(os.path.join(HERE, "configs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: float.fromhex(v["objective"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
