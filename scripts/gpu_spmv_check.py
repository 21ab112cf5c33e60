"""GPU check + tuning sweep of the SpMV tile kernel (run under gpurun)."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easylp_b200 import _lib as L
from oracle import gen

def rand_csr(m, n, lens, seed):
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens, dtype=np.int64)
    rp = np.zeros(m + 1, np.int64); np.cumsum(lens, out=rp[1:])
    nnz = int(rp[-1])
    ci = rng.integers(0, n, nnz).astype(np.int32)
    v = rng.normal(size=nnz)
    return rp.astype(np.int32), ci, v

def check(m, n, lens, seed):
    rp, ci, v = rand_csr(m, n, lens, seed)
    x = np.random.default_rng(seed + 1).normal(size=n)
    out = L.spmv(m, n, rp, ci, v, x)
    ref = np.zeros(m)
    for i in range(m):
        s = 0.0
        for k in range(rp[i], rp[i + 1]):
            s += v[k] * x[ci[k]]
        ref[i] = s
    exact = out.tobytes() == ref.tobytes()
    lanes = int(os.environ.get("ELP_SPMV_L", "1"))
    ok = exact if lanes == 1 else bool(np.allclose(out, ref, rtol=1e-12, atol=1e-12))
    print(f"spmv L={lanes} cap={os.environ.get('ELP_SPMV_CAP')} m={m} n={n} nnz={rp[-1]} bit-exact={exact} ok={ok} "
          f"maxdiff={np.abs(out-ref).max() if m else 0}", flush=True)
    return ok

allok = True
rng = np.random.default_rng(0)
for cap, lanes in (("0", "1"), ("64", "1"), ("16", "1"), ("0", "2"), ("0", "4"), ("64", "8"), ("0", "8"), ("32", "4"), ("16", "2")):
    os.environ["ELP_SPMV_CAP"] = cap
    os.environ["ELP_SPMV_L"] = lanes
    os.environ["ELP_SPMV_RPL"] = "1" if cap == "64" else "2"
    allok &= check(1, 5, [3], 1)
    allok &= check(7, 9, [0, 2, 0, 0, 5, 1, 0], 2)
    allok &= check(300, 50, rng.integers(0, 14, 300), 3)
    allok &= check(1000, 2000, rng.integers(0, 40, 1000), 4)
    allok &= check(5, 3000, [0, 2500, 1, 0, 9000], 5)      # rows longer than a stage -> pieces
    allok &= check(513, 100, np.r_[rng.integers(0, 3, 512), 700], 6)
os.environ["ELP_SPMV_CAP"] = "0"
os.environ.pop("ELP_SPMV_L"); os.environ.pop("ELP_SPMV_RPL")
print("ALL SPMV OK" if allok else "SPMV MISMATCH", flush=True)
if not allok:
    sys.exit(1)
if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    p = gen.sparse_planted(int(2_000_000 * scale), seed=0)
    m, n, nnz = p["m"], p["n"], int(p["row_ptr"][-1])
    b_csc = 12 * nnz + 4 * (n + 1) + 8 * m + 56 * n
    b_csr = 12 * nnz + 4 * (m + 1) + 8 * n + 40 * m
    os.environ["ELP_SPMV_DEBUG"] = "1"
    for lanes, ctas, nst, capsig in ((0, 0, 1, 100), (0, 0, 1, 100), (1, 0, 1, 100), (2, 0, 1, 100), (0, 5, 1, 100), (0, 4, 1, 100)):
        carve = 0
        os.environ["ELP_SPMV_RPL"] = str(nst)       # third field: rows per lane
        cw, capmul = lanes, capsig
        os.environ.update(ELP_SPMV_L=str(lanes), ELP_SPMV_CTAS=str(ctas), ELP_SPMV_NST=str(nst),
                          ELP_SPMV_CAPSIG_PCT=str(capsig))
        try:
            h = L.Pdlp(m, n, p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"], p["c"], p["lb"], p["ub"],
                       options=L.default_options(method=L.METHOD_PDLP, ruiz_iters=2))
            a, b = h.probe_step(30)
            c1, c2 = h.probe_spmv(30)
            h.close()
        except L.ElpError as e:
            print(json.dumps(dict(L=cw, ctas=ctas, nst=nst, capsig=capmul, carve=carve, error=str(e)[:200])), flush=True)
            continue
        print(json.dumps(dict(L=cw, ctas=ctas, nst=nst, capsig=capmul, carve=carve, primal_ms=round(a, 4), dual_ms=round(b, 4),
                              primal_gbs=round(b_csc / a / 1e6), dual_gbs=round(b_csr / b / 1e6), csr_ms=round(c1, 4),
                              csc_ms=round(c2, 4))), flush=True)
