"""ctypes front end of oracle/_build/libelp_oracle.so (the C restatements).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libelp_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(["make", "-C", _HERE])
        _lib = C.CDLL(_PATH)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def simplex_batch(A, b, c, lb=None, ub=None, sense=None, maximize=False, nthreads=1):
    A = np.ascontiguousarray(A, np.float64)
    B, m, n = A.shape
    b = np.ascontiguousarray(np.broadcast_to(b, (B, m)), np.float64)
    c = np.ascontiguousarray(np.broadcast_to(c, (B, n)), np.float64)
    lb = None if lb is None else np.ascontiguousarray(np.broadcast_to(lb, (B, n)), np.float64)
    ub = None if ub is None else np.ascontiguousarray(np.broadcast_to(ub, (B, n)), np.float64)
    sense = None if sense is None else np.ascontiguousarray(np.broadcast_to(sense, (B, m)), np.int8)
    status = np.zeros(B, np.int32)
    obj = np.zeros(B)
    x = np.zeros((B, n))
    piv = np.zeros(B, np.int32)
    f = lib().elpo_simplex_batch
    f.argtypes = [C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int]
    f.restype = None
    f(B, m, n, _p(A), _p(b), _p(c), _p(lb), _p(ub), _p(sense), int(maximize), _p(status), _p(obj), _p(x), _p(piv), nthreads)
    return status, obj, x, piv


def simplex_csr(m, n, row_ptr, col_idx, vals, sense, rhs, c, lb, ub, maximize=False):
    f = lib().elpo_simplex_csr
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 6
    f.restype = C.c_int
    row_ptr = np.ascontiguousarray(row_ptr, np.int32)
    if row_ptr.size == 0:
        row_ptr = np.zeros(1, np.int32)
    col_idx = np.ascontiguousarray(col_idx, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    sense = np.ascontiguousarray(sense, np.int8)
    rhs = np.ascontiguousarray(rhs, np.float64)
    c = np.ascontiguousarray(c, np.float64)
    lb = np.ascontiguousarray(np.broadcast_to(lb, (n,)), np.float64)
    ub = np.ascontiguousarray(np.broadcast_to(ub, (n,)), np.float64)
    x = np.zeros(n)
    y = np.zeros(max(m, 1))
    obj = C.c_double()
    piv = C.c_int()
    st = f(m, n, _p(row_ptr), _p(col_idx), _p(vals), _p(sense), _p(rhs), _p(c), int(maximize), _p(lb), _p(ub),
           C.addressof(obj), _p(x), _p(y), C.addressof(piv))
    return st, obj.value, x, y[:m], piv.value


def pdlp(p, eps=1e-6, max_iter=0, check_every=64, nthreads=0):
    """Runs oracle/pdlp_ref.c on a problem dict (keys of oracle/gen.py).  Returns (status, out[8], x, y);
    out = objective, iterations, restarts, rel primal res, rel dual res, rel gap, loop seconds, setup seconds."""
    f = lib().elpo_pdlp
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 2 + \
                 [C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3
    f.restype = C.c_int
    m, n = int(p["m"]), int(p["n"])
    row_ptr = np.ascontiguousarray(p["row_ptr"], np.int32)
    if row_ptr.size == 0:
        row_ptr = np.zeros(1, np.int32)
    col_idx = np.ascontiguousarray(p["col_idx"], np.int32)
    vals = np.ascontiguousarray(p["vals"], np.float64)
    sense = np.ascontiguousarray(p["sense"], np.int8)
    rhs = np.ascontiguousarray(p["rhs"], np.float64)
    c = np.ascontiguousarray(p["c"], np.float64)
    lb = np.ascontiguousarray(np.broadcast_to(p["lb"], (n,)), np.float64)
    ub = np.ascontiguousarray(np.broadcast_to(p["ub"], (n,)), np.float64)
    x = np.zeros(n)
    y = np.zeros(max(m, 1))
    out = np.zeros(8)
    st = f(m, n, _p(row_ptr), _p(col_idx), _p(vals), _p(sense), _p(rhs), _p(c), int(bool(p.get("maximize", False))),
           _p(lb), _p(ub), float(eps), int(max_iter), int(check_every), int(nthreads), _p(x), _p(y), _p(out))
    return st, out, x, y[:m]
