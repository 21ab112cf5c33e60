"""GPU: the reference's own models through the product path end to end — host DSL (easylp_b200.model) ->
elp_assemble_csr -> elp_solve_lp / elp_check_feasible on the B200 — against the committed goldens
(tests/golden/models.json: dense-oracle CSR bit-exact; status exact; objective <= 1e-6 rel) and the values the
reference's tests pin (README.md:28-39, test-DOP.R:53, test-unbounded.R:8-9, test-modified.R:17-41)."""
import warnings

import numpy as np
import pytest

import models
from easylp_b200 import _lib as L
from easylp_b200 import model as M
from fixtures import load_golden

pytestmark = pytest.mark.gpu
GOLD = load_golden()


def _build(name):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return models.ALL[name](M)


@pytest.mark.parametrize("name", sorted(models.ALL))
def test_device_assembly_bit_exact(name):
    lp = _build(name)
    g = GOLD[name]
    rp, ci, v = lp._csr()
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(ci, g["col_idx"])
    assert np.asarray(v).tobytes() == g["vals"].tobytes()
    assert lp.objective_fun.tobytes() == g["c"].tobytes()


CONTINUOUS = sorted(k for k, v in GOLD.items() if "highs" in v)


@pytest.mark.parametrize("method", ["auto", "pdlp", "pdlp-scatter"])
@pytest.mark.parametrize("name", CONTINUOUS)
def test_solve_matches_goldens(name, method):
    lp = _build(name)
    g = GOLD[name]
    if method == "pdlp-scatter":     # A'y by fp64 reductions from the dual kernel; plain "pdlp" is the CSC gather
        lp.solve(gpu_method="pdlp", gpu_transpose="scatter")
    else:
        lp.solve(gpu_method=method)
    want = g["highs"]["status"]
    assert lp.status == L.status_string(want)
    if want == 0:
        ref = g["highs"]["objective"]
        assert abs(lp.objective_value_raw - ref) <= 1e-6 * max(1.0, abs(ref))
        st = lp.pointer
        if st.method_used == L.METHOD_PDLP:
            assert st.rel_primal_res <= 1e-6 and st.rel_dual_res <= 1e-6 and st.rel_gap <= 1e-6


def test_readme_output():
    # README.md:28-39 : status optimal, objective 2, x = 1, y = 1
    lp = _build("readme").solve()
    assert lp.status == "optimal"
    assert abs(lp.objective_value - 2.0) <= 1e-9
    assert abs(lp.solution["x"] - 1.0) <= 1e-9 and abs(lp.solution["y"] - 1.0) <= 1e-9


def test_dop_objective():
    # test-DOP.R:53 : expect_equal(lp$objective_value, 3985000 - 45000)   (testthat tolerance 1.5e-8)
    lp = _build("dop").solve()
    assert abs(lp.objective_value - 3940000.0) <= 1.5e-8 * 3940000.0


def test_unbounded_is_infinite():
    # test-unbounded.R:8-9
    lp = _build("unbounded").solve()
    assert lp.status == "unbounded"
    assert lp.solution["x"] == np.inf and lp.objective_value == np.inf


def test_infeasible_then_uncon():
    # vignettes/constraints.Rmd:313-334
    lp = _build("infeasible_mean").solve()
    assert lp.status == "unfeasible"
    lp.uncon("limit")
    lp.solve()
    assert lp.status == "optimal" and abs(lp.objective_value - 12.0) <= 1e-6 * 12


def test_modified_solution_satisfies_constraints():
    # test-modified.R:17-21
    lp = _build("modified").solve()
    x, y = lp.solution["x"], lp.solution["y"]
    assert np.allclose(x.sum(axis=1), x.sum(axis=0), atol=1e-7)
    assert np.allclose(y.mean(axis=2).ravel(order="F"), [2, 3, 4, 5], atol=1e-7)
    assert np.allclose(np.diag(x)[1:3], [1, 2], atol=1e-7)


def test_new_constraint_invalidates_solution_message():
    # test-cyingair.R:31-33 (the message of R/class.R:384-385), on a continuous model
    lp = _build("transport_vignette").solve()
    assert lp.status == "optimal"
    x = lp["x"]
    lp.con(cut=M.Sum(x) <= 1)
    assert lp.status == "unsolved" and any("are unfeasible" in m for m in lp.messages)


def test_write_lp_readme():
    # lp_solve LP-format export of the assembled model
    lp = _build("readme")
    text = lp.write_lp()
    assert "max: +1.0 x +1.0 y;" in text
    assert "+1.0 x +2.0 y <= 3.0;" in text
    assert "-3.0 x +1.0 y >= -2.0;" in text
    assert "free x, y;" in text


def test_write_lp_round_trip_dop(tmp_path):
    """every coefficient survives the text round trip bit for bit (repr of a double is exact)"""
    import re
    lp = _build("dop")
    text = lp.write_lp(tmp_path / "dop.lp")
    g = GOLD["dop"]
    rows = [ln for ln in text.splitlines() if re.match(r"^[A-Za-z_].*: ", ln) and not ln.startswith(("min", "max"))]
    assert len(rows) == g["m"]
    vals = []
    for ln in rows:
        body = ln.split(": ", 1)[1]
        lhs = re.split(r" (<=|>=|=) ", body)[0]
        vals += [float(sign + num) for sign, num in re.findall(r"([+-])([0-9.e+-]+) ", lhs + " ")]
    assert np.array(vals).tobytes() == g["vals"].tobytes()
