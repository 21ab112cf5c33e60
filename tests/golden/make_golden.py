"""Regenerates tests/golden/models.json — run from the repo root:  python tests/golden/make_golden.py

What the fixture holds, per model of tests/models.py (the reference's own test / README / vignette models):

  * the canonical form of what `$solve()` hands to lp_solve (/root/reference/R/class.R:260-274), produced by the
    DENSE restatement of the reference's R arithmetic (oracle/dsl_ref.py): CSR of `constraint$mat`, `dir`, `rhs`,
    `objective_fun`, `objective_add`, replicated bounds, row names.  Doubles are stored as C99 hex strings so the
    comparison in the tests is bit-exact;
  * for the continuous models, status / objective / (unique) solution from an independent CPU LP solver
    (scipy.optimize.linprog -> HiGHS dual simplex), labelled "highs": a stand-in for lp_solve, which is not in
    this image.  Values the reference's own tests pin (test-DOP.R:53, README.md:28-39, test-unbounded.R:8-9)
    are NOT taken from here; they are written literally in tests/test_oracle_golden.py.

The reference itself is R (no R in this image), so it cannot be imported to generate vectors; this script and
the fixture are the committed substitute ("Oracle").
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import models  # noqa: E402
from oracle import dsl_ref  # noqa: E402


def hexes(a):
    return [float(v).hex() for v in np.asarray(a, dtype=float).ravel()]


def highs(can):
    from scipy.optimize import linprog
    from scipy.sparse import csr_matrix
    m, n = can["m"], can["n"]
    if not np.any(can["c"] != 0):
        return None
    A = csr_matrix((can["vals"], can["col_idx"], can["row_ptr"]), shape=(m, n))
    sense = can["sense"]
    sign = -1.0 if can["maximize"] else 1.0
    le, ge, eq = sense == 0, sense == 1, sense == 2
    A_ub = None
    b_ub = None
    if le.any() or ge.any():
        from scipy.sparse import vstack
        A_ub = vstack([A[le], -A[ge]])
        b_ub = np.r_[can["rhs"][le], -can["rhs"][ge]]
    bounds = [(None if not np.isfinite(l) else l, None if not np.isfinite(u) else u) for l, u in zip(can["lb"], can["ub"])]
    r = linprog(sign * can["c"], A_ub=A_ub, b_ub=b_ub, A_eq=A[eq] if eq.any() else None,
                b_eq=can["rhs"][eq] if eq.any() else None, bounds=bounds, method="highs-ds")
    status = {0: 0, 2: 2, 3: 3}.get(r.status, -1)          # lp_solve codes: 0 optimal, 2 infeasible, 3 unbounded
    out = {"status": status}
    if status == 0:
        out["objective"] = float(sign * r.fun).hex()
        out["x"] = hexes(r.x)
    return out


def highs_milp(can):
    """integer / binary models: HiGHS branch and cut (scipy.optimize.milp), stand-in for lp_solve's branch and bound"""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import csr_matrix
    m, n = can["m"], can["n"]
    A = csr_matrix((can["vals"], can["col_idx"], can["row_ptr"]), shape=(m, n))
    sign = -1.0 if can["maximize"] else 1.0
    lo = np.where(can["sense"] == 0, -np.inf, can["rhs"])
    hi = np.where(can["sense"] == 1, np.inf, can["rhs"])
    r = milp(sign * can["c"], constraints=LinearConstraint(A, lo, hi) if m else None, integrality=can["is_integer"].astype(int),
             bounds=Bounds(can["lb"], can["ub"]))
    status = {0: 0, 2: 2, 3: 3}.get(r.status, -1)
    out = {"status": status}
    if status == 0:
        out["objective"] = float(sign * r.fun).hex()
        out["x"] = hexes(r.x)
    return out


def main():
    warnings.simplefilter("ignore")
    out = {}
    for name, build in models.ALL.items():
        lp = build(dsl_ref)
        can = lp.canonical()
        rec = dict(m=int(can["m"]), n=int(can["n"]), row_ptr=[int(v) for v in can["row_ptr"]],
                   col_idx=[int(v) for v in can["col_idx"]], vals=hexes(can["vals"]), dir=can["dir"],
                   rhs=hexes(can["rhs"]), c=hexes(can["c"]), objective_add=float(can["objective_add"]).hex(),
                   lb=hexes(can["lb"]), ub=hexes(can["ub"]), maximize=bool(can["maximize"]),
                   names=can["names"], rownames=can["rownames"],
                   integer=bool(any(v.integer or v.binary for v in lp.variables.values())))
        rec["is_integer"] = [int(v) for v in can["is_integer"]]
        h = highs(can) if not rec["integer"] else highs_milp(can)
        if h is not None:
            rec["highs_milp" if rec["integer"] else "highs"] = h
        out[name] = rec
    with open(os.path.join(HERE, "models.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", os.path.join(HERE, "models.json"), {k: (v["m"], v["n"], len(v["vals"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
