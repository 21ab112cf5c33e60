"""Sensitivity ranging: `$sensitivity_objective` / `$sensitivity_rhs` of /root/reference/R/class.R:613-646
call lpSolveAPI::get.sensitivity.*.  CPU: the numpy restatement (oracle/ranging_ref.py) against HiGHS' own ranging on
non-degenerate models — the pin, since lp_solve is not in this image and the reference's tests print no sensitivity
values.  GPU: elp_sensitivity (simplex kernel's final basis + host ranging, csrc/sensitivity.cu) against the restatement."""
import numpy as np
import pytest

import models
from easylp_b200 import _lib as L
from easylp_b200 import model as M
from oracle import dsl_ref, ranging_ref


def brass():
    """vignettes/constraints.Rmd:208-220: max 8x + 6y; copper, zinc, silicon rows"""
    return dict(m=3, n=2, row_ptr=np.array([0, 2, 4, 5], np.int32), col_idx=np.array([0, 1, 0, 1, 1], np.int32),
                vals=np.array([0.90, 0.64, 0.10, 0.14, 0.04]), sense=np.zeros(3, np.int8), rhs=np.array([120.0, 15.0, 2.0]),
                c=np.array([8.0, 6.0]), lb=np.zeros(2), ub=np.full(2, np.inf), maximize=True)


def random_lp(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(3, 9)), int(rng.integers(2, 7))
    A = rng.uniform(0.1, 3.0, size=(m, n)) * (rng.random((m, n)) < 0.8)
    A[:, 0] = rng.uniform(0.5, 2.0, size=m)
    x0 = rng.uniform(0.5, 3.0, size=n)
    rhs = A @ x0 + rng.uniform(0.1, 2.0, size=m)
    rp = np.zeros(m + 1, np.int32); ci = []; v = []
    for i in range(m):
        nz = np.nonzero(A[i])[0]; ci += nz.tolist(); v += A[i, nz].tolist(); rp[i + 1] = len(ci)
    return dict(m=m, n=n, row_ptr=rp, col_idx=np.array(ci, np.int32), vals=np.array(v), sense=np.zeros(m, np.int8), rhs=rhs,
                c=rng.uniform(0.5, 4.0, size=n), lb=np.zeros(n), ub=np.full(n, 20.0), maximize=True)


def _nondegenerate(p, h):
    """strictly inside the bounds for every basic column / slack, non-zero reduced cost elsewhere"""
    x = h["x"]
    A = np.zeros((p["m"], p["n"]))
    for i in range(p["m"]):
        A[i, p["col_idx"][p["row_ptr"][i]:p["row_ptr"][i + 1]]] = p["vals"][p["row_ptr"][i]:p["row_ptr"][i + 1]]
    slack = p["rhs"] - A @ x
    ok = np.all((x[h["col_status"] == 0] > p["lb"][h["col_status"] == 0] + 1e-7) & (x[h["col_status"] == 0] < p["ub"][h["col_status"] == 0] - 1e-7))
    ok = ok and np.all(np.abs(slack[h["row_status"] == 0]) > 1e-7)
    return bool(ok) and int((h["col_status"] == 0).sum() + (h["row_status"] == 0).sum()) == p["m"]


def _close(a, b, tol=1e-7):
    a, b = np.asarray(a, float), np.asarray(b, float)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    return np.all(both_inf | (np.abs(a - b) <= tol * (1 + np.abs(b))))


def test_restatement_on_the_vignette_model_by_hand():
    # optimum x = 3600/31, y = 750/31 (copper and zinc binding, silicon slack)
    p = brass()
    h = ranging_ref.highs(p)
    if h is None:
        pytest.skip("scipy's HiGHS core not importable")
    of, ot, rf, rt = ranging_ref.ranging(p, h["x"], h["col_status"] == 0, h["row_status"] == 0)
    # cost of x may move while the slopes stay between the two binding rows: 6 * 0.10/0.14 <= cx <= 6 * 0.90/0.64
    assert _close([of[0], ot[0]], [6 * 0.10 / 0.14, 6 * 0.90 / 0.64])
    assert _close([of[1], ot[1]], [8 * 0.64 / 0.90, 8 * 0.14 / 0.10])
    # silicon is not binding: its rhs may fall to its activity 0.04 y and rise without limit
    assert _close(rf[2], 0.04 * 750 / 31) and rt[2] == np.inf
    assert _close([rf[0], rt[0]], [104.0, 135.0]) and _close([rf[1], rt[1]], [40 / 3, 151 / 9])


def test_restatement_against_highs_ranging():
    pinned = 0
    for p in [brass()] + [random_lp(s) for s in range(60)]:
        h = ranging_ref.highs(p)
        if h is None:
            pytest.skip("scipy's HiGHS core not importable")
        if not _nondegenerate(p, h):
            continue
        of, ot, rf, rt = ranging_ref.ranging(p, h["x"], h["col_status"] == 0, h["row_status"] == 0)
        bc = h["col_status"] == 0                       # HiGHS and the textbook agree on basic columns ...
        assert _close(of[bc], h["cost_dn"][bc]) and _close(ot[bc], h["cost_up"][bc]), (of, ot, h)
        br = h["row_status"] != 0                       # ... and on binding rows
        assert _close(rf[br], h["row_dn"][br]) and _close(rt[br], h["row_up"][br]), (rf, rt, h)
        pinned += 1
    assert pinned >= 30


@pytest.mark.gpu
def test_gpu_sensitivity_against_the_restatement():
    checked = 0
    for p in [brass()] + [random_lp(s) for s in range(60)]:
        h = ranging_ref.highs(p)
        if h is None or not _nondegenerate(p, h):
            continue
        st, obj, x, of, ot, rf, rt, du = L.sensitivity(p["m"], p["n"], p["row_ptr"], p["col_idx"], p["vals"], p["sense"], p["rhs"],
                                                       p["c"], p["lb"], p["ub"], maximize=p["maximize"])
        assert st == 0 and np.allclose(x, h["x"], atol=1e-8)
        rof, rot, rrf, rrt = ranging_ref.ranging(p, h["x"], h["col_status"] == 0, h["row_status"] == 0)
        assert _close(of, rof) and _close(ot, rot) and _close(rf, rrf) and _close(rt, rrt), (p, of, rof, ot, rot)
        checked += 1
    assert checked >= 30


@pytest.mark.gpu
def test_gpu_sensitivity_through_the_model_api():
    lp = M.easylp()
    x = lp.var("x", lower=0)
    y = lp.var("y", lower=0)
    lp.max(8 * x + 6 * y)
    lp.con(copper=0.90 * x + 0.64 * y <= 120, zinc=0.10 * x + 0.14 * y <= 15, silicon=0.04 * y <= 2)
    with pytest.raises(M.EasyLpError, match="not optimal"):
        lp.sensitivity_rhs
    lp.solve()
    r = lp.sensitivity_rhs                               # vignettes/constraints.Rmd:220 `lp$sensitivity_rhs |> round()`
    assert r.shape == (3, 3) and np.array_equal(np.round(r[:2]), [[104, 120, 135], [13, 15, 17]])
    assert r[2, 2] == np.inf and round(r[2, 0], 6) == round(0.04 * 750 / 31, 6)
    o = lp.sensitivity_objective
    assert o.shape == (2, 3) and np.allclose(o[:, 1], [8, 6])
    assert _close(o[0, [0, 2]], [6 * 0.10 / 0.14, 6 * 0.90 / 0.64])
    lp2 = models.ALL["investments_assembly"](M)
    lp2._stat = "optimal"
    with pytest.raises(M.EasyLpError, match="integer/binary"):
        lp2.sensitivity_objective
