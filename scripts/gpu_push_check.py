"""GPU probe (N >= 2): compact ghost exchange with and without push mode (pusher CTAs inside the SpMV grids) in ONE process.
The arithmetic is the same, so iterations and objective must be identical; only us/iteration may differ.
    python scripts/gpu_push_check.py N [--workloads pdlp,mcnf] [--scale S] [--variants "0;1;1,SEG=64"]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from easylp_b200 import _lib as L
from oracle import gen

ap = argparse.ArgumentParser()
ap.add_argument("n", type=int)
ap.add_argument("--workloads", default="pdlp,mcnf")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--variants", default="0;1")
ap.add_argument("--dense", default="0")
a = ap.parse_args()
for w in a.workloads.split(","):
    p = gen.sparse_planted(int(2_000_000 * a.scale), seed=0) if w == "pdlp" else gen.mcnf(K=max(1, int(50 * a.scale)))
    ref = p.get("obj_opt")
    for var in a.variants.split(";"):
        parts = var.split(",")
        env = {"ELP_GHOST_PUSH": parts[0], "ELP_PDLP_GHOST_DENSE": a.dense}
        for kv in parts[1:]:
            k, v = kv.split("=")
            env["ELP_GHOST_PUSH_" + k if k in ("SEG", "CTAS") else k] = v
        for k in [k for k in os.environ if k.startswith("ELP_GHOST_PUSH") or k.startswith("ELP_SPMV")]:
            del os.environ[k]
        os.environ.update(env)
        for rep in range(2):
            t0 = time.perf_counter()
            r = L.solve_lp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
                           options=L.default_options(method=L.METHOD_PDLP, devices=a.n))
            s = r.stats
            print(json.dumps({"w": w, "N": a.n, "env": env, "rep": rep, "status": r.status_string, "obj": r.objval,
                              "relerr": None if ref is None else abs(r.objval - ref) / max(1, abs(ref)),
                              "iters": s.iterations, "solve_ms": round(s.solve_ms, 1), "setup_ms": round(s.setup_ms, 1),
                              "wall_s": round(time.perf_counter() - t0, 2),
                              "us_per_iter": round(1e3 * s.solve_ms / max(s.iterations, 1), 2)}), flush=True)
