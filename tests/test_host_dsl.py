"""CPU: the product's host DSL (easylp_b200.model — sparse term lists) against the dense restatement of the
reference's R arithmetic (oracle/dsl_ref.py) and the committed goldens.  The device assembly itself
(elp_assemble_csr) is exercised by the -m gpu tests; here the term lists the host would hand to it are folded
by a scalar test-side loop (fixtures.ordered_fold, the written spec of the kernel) and compared BIT-EXACTLY.

The error cases mirror /root/reference/tests/testthat/test-constraints.R:22-35, test-aliases.R:24-26.
"""
import warnings

import numpy as np
import pytest

import models
from easylp_b200 import model as M
from easylp_b200 import lower
from fixtures import load_golden, model_fold, ordered_fold

GOLD = load_golden()


def _build(name, lowering=True):
    old = M.LOWERING
    M.LOWERING = lowering
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return models.ALL[name](M)
    finally:
        M.LOWERING = old


@pytest.mark.parametrize("lowering", [False, True])
@pytest.mark.parametrize("name", sorted(models.ALL))
def test_term_lists_fold_to_the_reference_matrix(name, lowering):
    lp = _build(name, lowering)
    g = GOLD[name]
    if not lowering:       # with lowering on, traceable for / sum_for rows are index-set families (lower.py)
        assert not any(isinstance(b, lower.LoweredCon) for b in lp._blocks)
    rp, ci, v, m = model_fold(lp)
    assert (m, lp.nvar) == (g["m"], g["n"])
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(ci, g["col_idx"])
    assert v.tobytes() == g["vals"].tobytes()
    assert list(lp.constraint.dir) == g["dir"]
    assert np.asarray(lp.constraint.rhs, float).tobytes() == g["rhs"].tobytes()
    assert lp.constraint.rownames == g["rownames"]
    # objective row: pending terms folded the same way
    oc, ov = lp._obj_terms
    c = np.zeros(lp.nvar)
    if oc.size:
        _, ci2, v2 = ordered_fold(np.zeros(oc.size, np.int64), oc, ov, 1, lp.nvar)
        c[ci2] = v2
    assert c.tobytes() == g["c"].tobytes()
    assert lp.objective_add == g["objective_add"]
    lb, ub = lp._bounds()
    assert lb.tobytes() == g["lb"].tobytes() and ub.tobytes() == g["ub"].tobytes()
    assert (lp.direction == "max") == g["maximize"]


def test_some_reference_models_are_lowered():
    """the reference's own models (test-DOP.R, test-aliases.R, test-forsplit.R, test-constraints.R, the vignettes'
    transport problem) write rows the trace accepts: `sum(x[f, ])`, `sum_for`, alias rows `ext[m]`; vector atoms
    (`x[, b, 2] >= 1`) and data-dependent ranges (`for (q in (p+1):n)`) keep the per-atom evaluation"""
    blocks = {n: _build(n, True)._blocks for n in models.ALL}
    lowered = {n for n, b in blocks.items() if any(isinstance(x, lower.LoweredCon) for x in b)}
    assert {"dop", "aliases", "forsplit", "constraints", "transport_vignette", "transport_sum_for",
            "investments_assembly"} <= lowered, lowered
    assert all(isinstance(x, lower.LoweredCon) for x in blocks["dop"])


def test_pickle_and_clone_drop_the_device_handle():
    """the device-resident model is a rebuildable cache — `$clone()` and saveRDS()/pickle never carry it"""
    import pickle

    class FakeHandle:                       # stands for _lib.Model: not picklable, must not be copied
        def __reduce__(self):
            raise TypeError("a device pointer cannot be pickled")

        def __deepcopy__(self, memo):
            return None

        def close(self):
            pass

    lp = _build("readme")
    lp._model = FakeHandle()
    back = pickle.loads(pickle.dumps(lp))
    assert back._model is None and back.constraint.rownames == lp.constraint.rownames
    assert back.constraint.rhs.tobytes() == lp.constraint.rhs.tobytes()
    assert lp.clone()._model is None and lp._model is not None
    lp._model = None


def _constraints_lp():
    lp = _build("constraints")
    return lp, lp["x"], lp["y"]


def test_invalid_variable_operations():
    # test-constraints.R:22-28
    lp, x, y = _constraints_lp()
    with pytest.raises(M.EasyLpError):
        lp.con(lambda: 2 / x[1, 1, 1] >= 0)
    with pytest.raises(M.EasyLpError):
        lp.con(lambda: x[1] * y[1] >= 0)
    with pytest.raises(M.EasyLpError):
        lp.con(lambda: abs(x) >= 2)
    with pytest.raises(M.EasyLpError):
        lp.con(lambda: y[9] >= 0)
    with pytest.raises(M.EasyLpError):
        lp.con(lambda: y[1, 1] >= 0)


def test_badly_defined_constraints():
    # test-constraints.R:30-35
    lp, x, y = _constraints_lp()
    with pytest.raises(M.EasyLpError):
        lp.con(5)
    with pytest.raises(M.EasyLpError):
        lp.con(lambda: M.rowSums(x == 1))
    with pytest.raises(M.EasyLpError):
        lp.con(lambda: x[0] == 1)
    with pytest.raises(M.EasyLpError):
        lp.con(lambda: x[1, 100, 1] == 0)
    m_before = lp.ncon
    assert m_before == GOLD["constraints"]["m"]          # failed $con calls leave the model untouched


def test_bad_alias_index():
    # test-aliases.R:24-26
    lp = _build("aliases")
    with pytest.raises(M.EasyLpError):
        lp.alias(err=lp["t"][1, 2, 3])


def test_pow_and_neq_are_errors():
    # R/methods.R:150-151, 216-217
    lp = M.easylp()
    x = lp.var("x", [1, 2])
    with pytest.raises(M.EasyLpError):
        x ** 2
    with pytest.raises(M.EasyLpError):
        x != 1


def test_solve_prechecks_keep_the_reference_messages():
    # R/class.R:253-258
    lp = M.easylp()
    with pytest.raises(M.EasyLpError, match="no variables"):
        lp.solve()
    lp.var("x")
    with pytest.raises(M.EasyLpError, match="objective function"):
        lp.solve()


def test_integer_variables_reach_the_branch_and_bound_entry_point():
    # R/class.R:264-276: set.type + lp_solve's branch and bound; here elp_solve_mip (csrc/mip.cu).  Without a CUDA device
    # the call fails loudly (no CPU fallback) instead of relaxing silently; the GPU run is tests/test_mip.py
    from easylp_b200 import _lib
    lp = _build("investments_assembly")
    if _lib.device_count() == 0:
        with pytest.raises(_lib.ElpError, match="no CUDA device"):
            lp.solve()


def test_division_is_multiplication_by_the_reciprocal():
    # R/methods.R:163 — x/3 stores 1/3 (0x1.5555555555555p-2), and (x*0.1)*3 is NOT folded to x*0.3
    lp = M.easylp()
    x = lp.var("x")
    a = (x / 3)
    assert a.t_val[0] == 1.0 / 3.0
    b = (x * 0.1) * 3
    assert b.t_val[0] == (0.1 * 3) and b.t_val[0] != 0.3


def test_large_to_infinity():
    # R/utils.R:172-176
    v = M.large_to_infinity(np.array([1e30, -1e30, 9.9e29, 0.0]))
    assert v.tolist() == [np.inf, -np.inf, 9.9e29, 0.0]
