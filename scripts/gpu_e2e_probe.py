import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen
p = gen.sparse_planted(2_000_000, seed=0)
opt = L.default_options(method=L.METHOD_PDLP, eps_rel=1e-6, max_iter=400000)
for i in range(3):
    t0 = time.perf_counter()
    h = L.Pdlp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"], options=opt)
    t1 = time.perf_counter()
    st = h.run()
    t2 = time.perf_counter()
    h.solution()
    t3 = time.perf_counter()
    h.close()
    t4 = time.perf_counter()
    print(json.dumps(dict(create=t1-t0, run_wall=t2-t1, run_dev_ms=st.solve_ms, iters=st.iterations, solution=t3-t2, close=t4-t3)), flush=True)
    if i == 0:
        # second solve on the same handle
        pass
