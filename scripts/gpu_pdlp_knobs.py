"""Full solves of a BASELINE config under ELP_* environment knobs (A/B runs of solver parameters).
usage: gpu_pdlp_knobs.py c4|c5|c2 "ELP_X=1;ELP_X=2,ELP_Y=3" [scale]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen
which = sys.argv[1]
configs = [dict(kv.split("=") for kv in c.split(",") if kv) for c in sys.argv[2].split(";")]
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
if which == "c4":
    p = gen.sparse_planted(int(2_000_000 * scale), seed=int(os.environ.get("SEED", "0")))
elif which == "c5":
    p = gen.mcnf(K=max(1, int(50 * scale)))
else:
    p = gen.transport(300, 300, seed=0)
m, n = p["m"], p["n"]
for cfg in configs:
    os.environ.update(cfg)
    kw = {}
    if "CHECK_EVERY" in cfg:
        kw["check_every"] = int(cfg["CHECK_EVERY"])
    h = L.Pdlp(m, n, p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
               maximize=p.get("maximize", False), options=L.default_options(method=L.METHOD_PDLP, max_iter=400000, **kw))
    st = h.run()
    x, y, obj = h.solution()
    h.close()
    for k in cfg:
        os.environ.pop(k, None)
    print(json.dumps(dict(which=which, cfg=cfg, status=st.status, iters=st.iterations, restarts=st.restarts, solve_ms=st.solve_ms,
                          ms_per_iter=st.solve_ms / max(st.iterations, 1), obj=obj, planted=p.get("obj_opt"),
                          pres=st.rel_primal_res, dres=st.rel_dual_res, gap=st.rel_gap)), flush=True)
