"""Probe the two fused iteration kernels on the C4 matrix under the ELP_SPMV_* settings in the environment."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
configs = [dict(os.environ)]
if len(sys.argv) > 3:
    configs = [dict(kv.split("=") for kv in c.split(",") if kv) for c in sys.argv[3].split(";")]
p = gen.sparse_planted(int(2_000_000 * scale), seed=0)
m, n, nnz = p["m"], p["n"], int(p["row_ptr"][-1])
b_csc = 12 * nnz + 4 * (n + 1) + 8 * m + 56 * n
b_csr = 12 * nnz + 4 * (m + 1) + 8 * n + 40 * m
for cfg in configs:
    env = {k: v for k, v in cfg.items() if k.startswith("ELP_")}
    os.environ.update(env)
    h = L.Pdlp(m, n, p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
               options=L.default_options(method=L.METHOD_PDLP, ruiz_iters=int(os.environ.get("RUIZ", "10"))))
    a, b = h.probe_step(reps)
    h.close()
    print(json.dumps(dict(env=env, primal_ms=a, dual_ms=b, primal_gbs=b_csc / a / 1e6, dual_gbs=b_csr / b / 1e6)), flush=True)
