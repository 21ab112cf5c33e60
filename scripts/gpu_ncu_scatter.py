"""ncu target: C4 in the scatter formulation without CUDA graphs, 130 iterations (direct launches only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = gen.sparse_planted(2_000_000, seed=0)
h = L.Pdlp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
           options=L.default_options(method=L.METHOD_PDLP, use_graph=0, transpose=mode))
st = h.run(130)
print("iterations", st.iterations, "transpose", h.transpose())
h.close()
