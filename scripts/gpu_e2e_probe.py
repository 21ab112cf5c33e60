"""create / run / solution / close wall times of the PDLP handle on C4 (or a scaled copy), 3 rounds per configuration.
usage: gpu_e2e_probe.py [scale] ["ELP_X=0;ELP_X=1"] [max_iter]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
configs = [dict(kv.split("=") for kv in c.split(",") if kv) for c in sys.argv[2].split(";")] if len(sys.argv) > 2 else [{}]
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 400000
p = gen.sparse_planted(int(2_000_000 * scale), seed=0)
opt = L.default_options(method=L.METHOD_PDLP, eps_rel=1e-6, max_iter=max_iter)
for cfg in configs:
    os.environ.update(cfg)
    for i in range(3):
        t0 = time.perf_counter()
        h = L.Pdlp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"], options=opt)
        t1 = time.perf_counter()
        st = h.run()
        t2 = time.perf_counter()
        x, y, obj = h.solution()
        t3 = time.perf_counter()
        h.close()
        t4 = time.perf_counter()
        print(json.dumps(dict(cfg=cfg, create=round(t1 - t0, 4), run_wall=round(t2 - t1, 4), iters=st.iterations, obj=obj,
                              solution=round(t3 - t2, 4), close=round(t4 - t3, 4))), flush=True)
    for k in cfg:
        os.environ.pop(k, None)
