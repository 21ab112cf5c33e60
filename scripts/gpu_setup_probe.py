"""Times elp_pdlp_create (H2D + transpose + scaling + power iteration) on the C4 problem."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen
p = gen.sparse_planted(int(2_000_000 * float(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000), seed=0)
for i in range(3):
    t = time.perf_counter()
    h = L.Pdlp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
               options=L.default_options(method=L.METHOD_PDLP))
    dt = time.perf_counter() - t
    t = time.perf_counter(); h.close(); dc = time.perf_counter() - t
    print(json.dumps(dict(create_s=dt, close_s=dc)), flush=True)
