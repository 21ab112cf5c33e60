"""ncu target: one BASELINE config in the gather formulation WITHOUT CUDA graphs (direct launches only), 130 iterations.
    python scripts/gpu_ncu_target.py pdlp|mcnf"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen
which = sys.argv[1] if len(sys.argv) > 1 else "pdlp"
p = gen.sparse_planted(2_000_000, seed=0) if which == "pdlp" else gen.mcnf(K=50)
h = L.Pdlp(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
           options=L.default_options(method=L.METHOD_PDLP, use_graph=0))
st = h.run(130)
print("iterations", st.iterations, "kernel launches", st.kernel_launches)
h.close()
