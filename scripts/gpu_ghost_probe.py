"""GPU probe (N >= 2): where does the time of the ghost exchange go?  Fixed iteration count per variant of
ELP_GHOST_DEBUG (pdlp.cu): 1 no remote stores, 2 no wait, 4 no release, 8 every warp polls.  Timing only: variants that cut
the data flow do not converge.   python scripts/gpu_ghost_probe.py N [--iters K]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from easylp_b200 import _lib as L
from oracle import gen

ap = argparse.ArgumentParser()
ap.add_argument("n", type=int)
ap.add_argument("--iters", type=int, default=4000)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--variants", default="0,8,1,6,7,p2p0")
a = ap.parse_args()
p = gen.sparse_planted(int(2_000_000 * a.scale), seed=0)
for v in a.variants.split(","):
    os.environ.pop("ELP_PDLP_P2P", None)
    os.environ.pop("ELP_SPMV_L2HINTS", None)
    os.environ.pop("ELP_GHOST_LOCAL_ONLY", None)
    if v.startswith("local"):                   # e.g. local1: K1 keeps x-bar at home, local2: K2 keeps y at home, local3: both
        os.environ["ELP_GHOST_LOCAL_ONLY"] = v[5:]
        v = "64"
    os.environ["ELP_GHOST_DEBUG"] = "0"
    if v.startswith("hint"):                    # e.g. hint19: ELP_SPMV_L2HINTS=19 (bit 16 = strided tile walk, timing only)
        os.environ["ELP_SPMV_L2HINTS"] = v[4:]
    elif v == "p2p0":
        os.environ["ELP_PDLP_P2P"] = "0"
    else:
        os.environ["ELP_GHOST_DEBUG"] = v
    for rep in range(2):
        r = L.solve_lp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
                       options=L.default_options(method=L.METHOD_PDLP, devices=a.n, max_iter=a.iters))
    s = r.stats
    print(json.dumps({"N": a.n, "variant": v, "iters": s.iterations, "us_per_iter": 1e3 * s.solve_ms / max(s.iterations, 1),
                      "status": r.status_string}), flush=True)
