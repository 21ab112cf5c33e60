"""GPU probe: where does elp_pdlp_destroy spend its time on short solves (config 2 / config 5)?  ELP_PDLP_DEBUG prints the
destructor's phases."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ELP_PDLP_DEBUG"] = "1"
from easylp_b200 import _lib as L
from oracle import gen
for name, p in (("c2", gen.transport(300, 300, seed=0)), ("c5", gen.mcnf(K=50))):
    for rep in range(3):
        t0 = time.perf_counter()
        h = L.Pdlp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
                   options=L.default_options(method=L.METHOD_PDLP))
        t1 = time.perf_counter()
        st = h.run()
        t2 = time.perf_counter()
        h.solution()
        t3 = time.perf_counter()
        h.close()
        t4 = time.perf_counter()
        print(f"{name} rep {rep}: create {t1-t0:.4f} run {t2-t1:.4f} ({st.iterations} it) solution {t3-t2:.4f} close {t4-t3:.4f}", flush=True)
