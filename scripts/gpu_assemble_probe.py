"""Times elp_assemble_csr on the C4 term stream (emission order = a seeded permutation of the matrix entries, plus
5 % duplicate terms), checks the result against the generator's CSR, prints algorithmic GB/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from easylp_b200 import _lib as L
from oracle import gen
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = gen.sparse_planted(int(2_000_000 * scale), seed=0)
m, n, nnz = p["m"], p["n"], int(p["row_ptr"][-1])
r, c, v = gen.term_stream(p, 0.05, 0)
T = r.size
for _ in range(reps):
    rp, ci, vv, st = L.assemble_csr(r, c, v, m, n)
ok = np.array_equal(rp, p["row_ptr"]) and np.array_equal(ci, p["col_idx"]) and vv.tobytes() == p["vals"].tobytes()
alg = 16 * T + 12 * nnz + 4 * (m + 1)
print(json.dumps(dict(T=int(T), m=m, n=n, nnz=nnz, bit_exact=bool(ok), device_ms=st.solve_ms, total_ms=st.total_ms,
                      launches=int(st.kernel_launches), alg_bytes=alg, gbs=alg / st.solve_ms / 1e6)))
# the same terms in `$con()` emission order (rows ascending, columns ascending): the two-pass ordered path
order = np.lexsort((c, r))
sr, sc, sv = r[order], c[order], v[order]
for _ in range(reps):
    rp, ci, vv, st = L.assemble_csr(sr, sc, sv, m, n)
ok = np.array_equal(rp, p["row_ptr"]) and np.array_equal(ci, p["col_idx"]) and vv.tobytes() == p["vals"].tobytes()
print(json.dumps(dict(stream="ordered", bit_exact=bool(ok), device_ms=st.solve_ms, launches=int(st.kernel_launches),
                      gbs=alg / st.solve_ms / 1e6)))
