"""A/B of the two PDLP formulations on the C4 matrix: per-kernel times (probe_step) and one full solve each.
usage: gpu_scatter_probe.py [scale] ["ELP_A=1,ELP_B=2;ELP_A=2"]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
configs = [dict(kv.split("=") for kv in c.split(",") if kv) for c in sys.argv[2].split(";")] if len(sys.argv) > 2 else [{}]
solve = int(os.environ.get("SOLVE", "1"))
p = gen.sparse_planted(int(2_000_000 * scale), seed=0)
m, n, nnz = p["m"], p["n"], int(p["row_ptr"][-1])
for cfg in configs:
    os.environ.update(cfg)
    h = L.Pdlp(m, n, p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
               options=L.default_options(method=L.METHOD_PDLP))
    a, b = h.probe_step(50)
    out = dict(cfg=cfg, transpose=h.transpose(), primal_ms=a, dual_ms=b, pair_ms=a + b)
    if solve:
        h.reset()
        st = h.run()
        x, y, obj = h.solution()
        out.update(status=st.status, iters=st.iterations, solve_ms=st.solve_ms, iter_per_s=st.iterations / st.solve_ms * 1e3,
                   ms_per_iter=st.solve_ms / st.iterations, obj=obj, planted=p["obj_opt"], pres=st.rel_primal_res,
                   dres=st.rel_dual_res, gap=st.rel_gap, restarts=st.restarts)
    h.close()
    for k in cfg:
        os.environ.pop(k, None)
    print(json.dumps(out), flush=True)
