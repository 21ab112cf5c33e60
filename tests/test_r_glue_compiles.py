"""CPU: the `.Call` glue of the R package (rpkg/src/r_glue.c) type-checks against include/easylp_abi.h.

R is not installed in this image (nor on the GPU box), so the glue can only be compiled against a minimal mock of R's
C API (tests/r_mock/); it is never linked or run.  This keeps the glue in step with the ABI: a changed signature in
include/easylp_abi.h breaks this test."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "rpkg", "src", "r_glue.c")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_glue_typechecks_against_the_abi():
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-fsyntax-only",
           "-I", os.path.join(ROOT, "tests", "r_mock"), "-I", os.path.join(ROOT, "include"), GLUE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_registered_routines_match_their_definitions():
    src = open(GLUE).read()
    table = dict((n, int(k)) for n, k in re.findall(r'\{"(easylp_\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', src))
    assert set(table) == {"easylp_assemble_csr", "easylp_assemble_lowered", "easylp_model_assemble", "easylp_model_csr",
                          "easylp_model_solve", "easylp_model_valid", "easylp_solve_lp", "easylp_solve_mip", "easylp_sensitivity", "easylp_sensitivity",
                          "easylp_check_feasible", "easylp_solve_batch", "easylp_device_count"}
    for name, nargs in table.items():
        sig = re.search(r"SEXP %s\(([^)]*)\)" % name, src).group(1)
        got = 0 if sig.strip() == "void" else sig.count("SEXP")
        assert got == nargs, (name, got, nargs)


def test_r_side_calls_only_registered_routines():
    """every .Call("easylp_...") in rpkg/R/gpu_solve.R names a routine of the registration table, with the right arity"""
    src = open(GLUE).read()
    table = dict((n, int(k)) for n, k in re.findall(r'\{"(easylp_\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', src))
    r_src = open(os.path.join(ROOT, "rpkg", "R", "gpu_solve.R")).read()
    calls = re.findall(r'\.Call\("(easylp_\w+)"', r_src)
    assert calls and set(calls) <= set(table), set(calls) - set(table)
    # arity: count top-level commas of each call's argument list
    for m in re.finditer(r'\.Call\("(easylp_\w+)"', r_src):
        i, depth, args = m.end(), 1, 0
        while depth:
            ch = r_src[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            args += ch == "," and depth == 1
            i += 1
        assert args == table[m.group(1)], (m.group(1), args, table[m.group(1)])
