"""Times the batched simplex kernel on the C3 batch (device-resident), for ncu captures and quick sweeps."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from easylp_b200 import _lib as L
from oracle import gen
B = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = gen.dense_batch(B=B, seed=0)
h = L.Batch(d["A"], d["b"], d["c"], d["lb"], d["ub"], d["sense"])
for _ in range(reps):
    st = h.run()
status, obj, x = h.fetch()
print(json.dumps(dict(B=B, ms=st.solve_ms, lps_per_s=B / st.solve_ms * 1e3, pivots=int(st.iterations), optimal=int((status == 0).sum()))))
h.close()
