"""CPU: pins the oracles against every golden the reference holds for the path.

The reference is R and cannot run in this image, so the anchors are (i) the values the reference's own tests /
README print, written literally below with their file:line, and (ii) an independent CPU solver (HiGHS via scipy,
labelled as a stand-in for lp_solve) recorded in tests/golden/models.json by tests/golden/make_golden.py.
"""
import warnings

import numpy as np
import pytest

import models
from fixtures import load_golden
from oracle import cbind, dsl_ref, pdlp_ref

GOLD = load_golden()


@pytest.mark.parametrize("name", sorted(models.ALL))
def test_fixture_is_what_the_dense_oracle_builds(name):
    """tests/golden/models.json is reproducible bit for bit from oracle/dsl_ref.py"""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        can = models.ALL[name](dsl_ref).canonical()
    g = GOLD[name]
    assert (can["m"], can["n"]) == (g["m"], g["n"])
    assert np.array_equal(can["row_ptr"], g["row_ptr"]) and np.array_equal(can["col_idx"], g["col_idx"])
    for k in ("vals", "rhs", "c", "lb", "ub"):
        assert np.asarray(can[k], float).tobytes() == g[k].tobytes(), k
    assert can["dir"] == g["dir"] and can["rownames"] == g["rownames"] and can["names"] == g["names"]
    assert can["objective_add"] == g["objective_add"] and bool(can["maximize"]) == g["maximize"]


def _dense(g):
    out = np.zeros((g["m"], g["n"]))
    for i in range(g["m"]):
        a, b = g["row_ptr"][i], g["row_ptr"][i + 1]
        out[i, g["col_idx"][a:b]] = g["vals"][a:b]
    return out


def test_g1_readme_rows():
    # /root/reference/README.md:16-24 ; `y >= 3*x - 2` moves left as -3x + y >= -2 (R/methods.R:202-214)
    g = GOLD["readme"]
    assert _dense(g).tolist() == [[1.0, 2.0], [-3.0, 1.0]]
    assert g["dir"] == ["<=", ">="] and g["rhs"].tolist() == [3.0, -2.0]
    assert g["c"].tolist() == [1.0, 1.0] and g["maximize"]
    assert np.all(np.isinf(g["lb"])) and np.all(np.isinf(g["ub"]))       # default bounds are free (R/class.R:86)


def test_g2_dop_structure():
    # /root/reference/tests/testthat/test-DOP.R ; SURVEY §8c G2
    g = GOLD["dop"]
    assert (g["n"], g["m"], g["vals"].size) == (14, 11, 36)
    assert g["c"].tolist() == [132, 138, 119, 132, 138, 131, 135, 134, 47, 58, 52, 56, 51, 59]
    assert g["objective_add"] == -45000.0
    assert g["dir"] == ["=="] * 2 + ["<="] * 6 + [">="] * 3
    assert g["rownames"][:3] == ["tdm_ext[m=A]", "tdm_ext[m=B]", "recolleccio[d=Empordà]"]


def test_g4_number_on_the_left():
    # /root/reference/vignettes/constraints.Rmd:225-230: `2 >= x`  ->  row [-1] >= -2
    g = GOLD["rhs_variable"]
    assert _dense(g).tolist() == [[-1.0]] and g["dir"] == [">="] and g["rhs"].tolist() == [-2.0]


def test_g6_structure_goldens():
    # test-constraints.R:11-20 — row counts per name after uncon("r3")
    g = GOLD["constraints"]
    base = [r.split("[")[0] for r in g["rownames"]]
    assert {b: base.count(b) for b in set(base)} == {"r1": 3, "r2": 6, "r4": 12, "r5": 3, "r6": 12}
    assert g["rownames"][3:6] == ["r2[a=1,b=1]", "r2[a=1,b=2]", "r2[a=1,b=3]"]
    assert g["dir"][-1] == ">"                                            # strict dirs are stored as written
    # cumsum(2*y + 1) >= 0 : lower-triangular 2s, rhs -1 -2 -3 (R/methods.R:228-242, :214)
    r5 = [i for i, b in enumerate(base) if b == "r5"]
    assert _dense(g)[r5][:, 12:15].tolist() == [[2, 0, 0], [2, 2, 0], [2, 2, 2]]
    assert g["rhs"][r5].tolist() == [-1.0, -2.0, -3.0]
    # test-forsplit.R:4-6 — triangular nested for: 4+3+2+1 rows named hi[i=..,j=..]
    f = GOLD["forsplit"]
    assert f["m"] == 10 and f["rownames"][0] == "hi[i=1,j=1]" and f["rownames"][-1] == "hi[i=4,j=4]"
    # test-aliases.R:18-21
    a = GOLD["aliases"]
    assert a["rownames"] == ["cap[i=A]", "cap[i=B]", "dem[j=1]", "dem[j=2]"]
    assert _dense(a).tolist() == [[1, 0, 1, 0], [0, 1, 0, 1], [1, 1, 0, 0], [0, 0, 1, 1]]


def test_sum_for_objective_equals_vectorised_objective():
    # vignettes/easylp.Rmd:135 — sum_for(f, m, cost[f,m]*x[f,m]) is the same cost vector as sum(cost * x)
    a, b = GOLD["transport_vignette"], GOLD["transport_sum_for"]
    assert a["c"].tobytes() == b["c"].tobytes() and a["vals"].tobytes() == b["vals"].tobytes()


# ---- solver oracles against the reference's pinned results and the independent solver ---------------
PINNED = {               # what the reference's own tests / README assert
    "readme": (0, 2.0, [1.0, 1.0]),            # README.md:28-39
    "dop": (0, 3985000.0, None),               # test-DOP.R:53  (3 985 000 - 45 000 with the addend)
    "unbounded": (3, np.inf, [np.inf]),        # test-unbounded.R:8-9
    "infeasible_mean": (2, None, None),        # vignettes/constraints.Rmd:313-334 (prose only)
    "transport_vignette": (0, 5785.0, None),   # HiGHS value; the vignette prints no number
}


def _solve_simplex(g):
    return cbind.simplex_csr(g["m"], g["n"], g["row_ptr"], g["col_idx"], g["vals"], g["sense"], g["rhs"], g["c"],
                             g["lb"], g["ub"], g["maximize"])


@pytest.mark.parametrize("name", sorted(k for k, v in GOLD.items() if "highs" in v))
def test_simplex_oracle_vs_goldens(name):
    g = GOLD[name]
    st, obj, x, y, piv = _solve_simplex(g)
    assert st == g["highs"]["status"]
    if st == 0:
        ref = g["highs"]["objective"]
        assert abs(obj - ref) <= 1e-9 * max(1.0, abs(ref))
    if name in PINNED:
        pst, pobj, px = PINNED[name]
        assert st == pst
        if pobj is not None:
            assert obj == pobj or abs(obj - pobj) <= 1e-9 * abs(pobj)
        if px is not None:
            assert np.allclose(x, px, atol=1e-9) or np.array_equal(x, px)


@pytest.mark.parametrize("name", sorted(k for k, v in GOLD.items() if "highs" in v))
def test_pdlp_oracle_vs_goldens(name):
    """oracle/pdlp_ref.py (the restatement of the GPU algorithm): status exact, objective <= 1e-6 rel, residuals <= 1e-6"""
    g = GOLD[name]
    r = pdlp_ref.solve(g["m"], g["n"], g["row_ptr"], g["col_idx"], g["vals"], g["sense"], g["rhs"], g["c"], g["lb"],
                       g["ub"], maximize=g["maximize"])
    assert r["status"] == g["highs"]["status"]
    if r["status"] == 0:
        ref = g["highs"]["objective"]
        assert abs(r["obj"] - ref) <= 1e-6 * max(1.0, abs(ref))
        assert r["pres"] <= 1e-6 and r["dres"] <= 1e-6 and r["gap"] <= 1e-6


@pytest.mark.parametrize("name", ["readme", "dop", "transport_vignette", "modified"])
def test_pdlp_c_port_matches_python_restatement(name):
    """oracle/pdlp_ref.c (the timed CPU baseline) follows the same iteration as oracle/pdlp_ref.py"""
    g = GOLD[name]
    st, out, x, y = cbind.pdlp(g, nthreads=1)
    ref = g["highs"]["objective"]
    assert st == 0 and abs(out[0] - ref) <= 1e-6 * max(1.0, abs(ref))
    assert out[3] <= 1e-6 and out[4] <= 1e-6 and out[5] <= 1e-6


def test_modified_solution_satisfies_its_constraints():
    # test-modified.R:17-21 — the reference only checks that the returned x satisfies the rows
    g = GOLD["modified"]
    st, obj, x, y, piv = _solve_simplex(g)
    assert st == 0
    lhs = _dense(g) @ x
    assert np.allclose(lhs, g["rhs"], atol=1e-9)
