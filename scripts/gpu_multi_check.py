"""GPU probe: config 4 (or a scaled copy) through elp_solve_lp(devices = N), timing and parity vs the planted optimum.
    python scripts/gpu_multi_check.py [N ...] [--scale S] [--workload pdlp|mcnf]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from easylp_b200 import _lib as L
from oracle import gen

ap = argparse.ArgumentParser()
ap.add_argument("n", nargs="*", type=int, default=[1, 2])
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--workload", default="pdlp")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
if a.workload == "pdlp":
    p = gen.sparse_planted(int(2_000_000 * a.scale), seed=0)
else:
    p = gen.mcnf(K=max(1, int(50 * a.scale)))
ref = p.get("obj_opt")
for N in a.n:
    for rep in range(a.reps):
        t0 = time.perf_counter()
        r = L.solve_lp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
                       options=L.default_options(method=L.METHOD_PDLP, devices=N))
        dt = time.perf_counter() - t0
        s = r.stats
        print(json.dumps({"N": N, "rep": rep, "status": r.status_string, "obj": r.objval, "ref": ref,
                          "relerr": None if ref is None else abs(r.objval - ref) / max(1, abs(ref)),
                          "iters": s.iterations, "restarts": s.restarts, "solve_ms": s.solve_ms, "setup_ms": s.setup_ms,
                          "wall_s": dt, "iter_per_s": s.iterations / (s.solve_ms * 1e-3), "us_per_iter": 1e3 * s.solve_ms / max(s.iterations, 1),
                          "gap": s.rel_gap, "pres": s.rel_primal_res, "dres": s.rel_dual_res}), flush=True)
