"""CPU: the C-ABI library loads and exports every symbol include/easylp_abi.h declares; without a CUDA device
the compute entry points fail loudly (rc != 0 + error text) instead of falling back to anything."""
import os
import re

import numpy as np
import pytest

from easylp_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "easylp_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(elp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = L.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(L.ABI_SYMBOLS) == declared          # the ctypes binding covers the whole header


def test_status_strings_are_the_reference_switch():
    # /root/reference/R/class.R:279-295
    want = {0: "optimal", 1: "sub-optimal", 2: "unfeasible", 3: "unbounded", 4: "degenerate model",
            5: "numerical failure encountered", 6: "process aborted", 7: "timeout"}
    for code, text in want.items():
        assert L.status_string(code) == text
    assert L.status_string(42) == "undocumented status"


def test_default_options():
    o = L.default_options()
    assert o.eps_rel == 1e-6 and o.method == L.METHOD_AUTO and o.check_every > 0


def _has_gpu():
    try:
        return L.device_count() > 0
    except L.ElpError:
        return False


@pytest.mark.skipif(_has_gpu(), reason="this test pins the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(L.ElpError):
        L.assemble_csr(np.zeros(1, np.int32), np.zeros(1, np.int32), np.ones(1), 1, 1)
    with pytest.raises(L.ElpError):
        L.solve_lp(0, 1, [0], [], [], [], [], [1.0], [0.0], [1.0])
    with pytest.raises(L.ElpError):
        L.solve_batch(np.ones((1, 1, 1)), np.ones((1, 1)), np.ones((1, 1)), None, None, None)
    from easylp_b200 import lower
    empty = lower.pack([])
    with pytest.raises(L.ElpError, match="no CUDA device"):
        L.assemble_lowered(np.zeros(1, np.int32), np.zeros(1, np.int32), np.ones(1), empty, 1, 1)
    with pytest.raises(L.ElpError, match="no CUDA device"):
        L.expand_terms(empty)
    with pytest.raises(L.ElpError, match="no CUDA device"):
        L.Model(np.zeros(1, np.int32), np.zeros(1, np.int32), np.ones(1), empty, 1, 1)


@pytest.mark.skipif(_has_gpu(), reason="this test pins the no-GPU behaviour")
def test_the_host_dsl_cannot_solve_without_the_device():
    """the product path through the R6 mirror: building terms is host work, folding and solving are not"""
    from easylp_b200 import model as M
    lp = M.easylp()
    x, y = lp.var("x"), lp.var("y")
    lp.max(x + y)
    lp.con(x + 2 * y <= 3, y >= 3 * x - 2)
    with pytest.raises(L.ElpError, match="no CUDA device"):
        lp.solve()
    with pytest.raises(L.ElpError, match="no CUDA device"):
        lp.constraint.mat
