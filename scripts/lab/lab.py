"""Algorithm lab (CPU, reduced sizes): runs scripts/lab/pdlp_lab.c with LAB_* env knobs. Not part of the product."""
import ctypes as C, os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import gen
lib = C.CDLL("/tmp/libpdlp_lab.so")
def _p(a): return a.ctypes.data_as(C.c_void_p)
def run(p, eps=1e-6, max_iter=200000):
    f = lib.elpo_pdlp_lab
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p]*6 + [C.c_int] + [C.c_void_p]*2 + [C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_void_p]*3
    f.restype = C.c_int
    m, n = int(p["m"]), int(p["n"])
    a = [np.ascontiguousarray(p[k], t) for k, t in (("row_ptr", np.int32), ("col_idx", np.int32), ("vals", np.float64), ("sense", np.int8), ("rhs", np.float64), ("c", np.float64))]
    lb = np.ascontiguousarray(np.broadcast_to(p["lb"], (n,)), np.float64); ub = np.ascontiguousarray(np.broadcast_to(p["ub"], (n,)), np.float64)
    x = np.zeros(n); y = np.zeros(max(m,1)); out = np.zeros(8)
    st = f(m, n, *[_p(v) for v in a], int(bool(p.get("maximize", False))), _p(lb), _p(ub), eps, max_iter, 64, 8, _p(x), _p(y), _p(out))
    return st, out
if __name__ == "__main__":
    which = sys.argv[1]
    if which == "c4": p = gen.sparse_planted(int(sys.argv[2]) if len(sys.argv) > 2 else 200000, seed=0)
    elif which == "c2": p = gen.transport(300, 300, seed=0)
    elif which == "c5": p = gen.mcnf(K=int(sys.argv[2]) if len(sys.argv) > 2 else 5)
    st, out = run(p)
    ref = p.get("obj_opt")
    print(json.dumps({"w": which, "status": st, "obj": out[0], "ref": ref, "relerr": (abs(out[0]-ref)/max(1,abs(ref)) if ref is not None else None), "iters": int(out[1]), "restarts": int(out[2]), "gap": out[5], "loop_s": round(out[6],2),
       "knobs": {k: v for k, v in os.environ.items() if k.startswith("LAB_")}}))
