"""Shared helpers of the test-suite: golden fixture loading and a test-side ordered fold."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SENSE = {"<=": 0, "<": 0, ">=": 1, ">": 1, "==": 2}


def unhex(a):
    return np.array([float.fromhex(s) for s in a], dtype=float)


def load_golden():
    with open(os.path.join(HERE, "golden", "models.json")) as f:
        raw = json.load(f)
    out = {}
    for name, v in raw.items():
        g = dict(v)
        for k in ("vals", "rhs", "c", "lb", "ub"):
            g[k] = unhex(v[k])
        g["row_ptr"] = np.array(v["row_ptr"], np.int32)
        g["col_idx"] = np.array(v["col_idx"], np.int32)
        g["objective_add"] = float.fromhex(v["objective_add"])
        g["sense"] = np.array([SENSE[d] for d in v["dir"]], np.int8)
        for key in ("highs", "highs_milp"):           # continuous models / models with integer columns (branch and cut)
            if key in v:
                h = dict(v[key])
                if "objective" in h:
                    h["objective"] = float.fromhex(h["objective"])
                    h["x"] = unhex(h["x"])
                g[key] = h
        g["is_integer"] = np.array(v.get("is_integer", [0] * v["n"]), np.uint8)
        out[name] = g
    return out


def ordered_fold(rows, cols, vals, m, n):
    """Canonical CSR of a term list: duplicates of one (row, col) summed strictly left to right in emission
    order, exact zeros dropped, row-major ascending column.  Scalar Python on purpose (the spec of
    elp_assemble_csr, include/easylp_abi.h)."""
    acc = {}
    for r, c, v in zip(np.asarray(rows).tolist(), np.asarray(cols).tolist(), np.asarray(vals, dtype=float).tolist()):
        k = (r, c)
        acc[k] = acc[k] + v if k in acc else v
    keys = sorted(k for k, s in acc.items() if s != 0.0)
    row_ptr = np.zeros(m + 1, np.int64)
    for r, _ in keys:
        row_ptr[r + 1] += 1
    return (np.cumsum(row_ptr).astype(np.int32), np.array([c for _, c in keys], np.int32),
            np.array([acc[k] for k in keys], dtype=float))


def model_fold(lp):
    """Canonical CSR of a product-side model on the CPU: the eager blocks' term lists plus the lowered blocks' families
    expanded by the plain-Python oracle (oracle/lower_ref.py), folded by its two-level ordered fold (group 0 = the
    plain left fold of ordered_fold).  Returns (row_ptr, col_idx, vals, m)."""
    from easylp_b200 import lower
    from oracle import lower_ref
    blocks = lp._blocks
    m = sum(b.nrow for b in blocks)
    offs = np.cumsum([0] + [b.nrow for b in blocks])
    eager = [(b, o) for b, o in zip(blocks, offs) if not isinstance(b, lower.LoweredCon)]
    low = [(b, int(o)) for b, o in zip(blocks, offs) if isinstance(b, lower.LoweredCon)]
    rows = np.concatenate([b.t_row + o for b, o in eager]) if eager else np.zeros(0, np.int64)
    cols = np.concatenate([b.t_col for b, _ in eager]) if eager else np.zeros(0, np.int64)
    vals = np.concatenate([b.t_val for b, _ in eager]) if eager else np.zeros(0)
    if not low:
        return ordered_fold(rows, cols, vals, m, lp.nvar) + (m,)
    packed = lower.pack(low)
    r2, c2, v2, g2 = lower_ref.expand(packed)
    return lower_ref.fold(np.r_[rows, r2], np.r_[cols, c2], np.r_[vals, v2], np.r_[np.zeros(rows.size, np.int32), g2],
                          packed, m) + (m,)


def model_terms(lp):
    """(rows, cols, vals, m) of a product-side model (easylp_b200.model.easylp) before device assembly."""
    blocks = lp._blocks
    m = sum(b.nrow for b in blocks)
    if m == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0), 0
    offs = np.cumsum([0] + [b.nrow for b in blocks])
    rows = np.concatenate([b.t_row + o for b, o in zip(blocks, offs)])
    cols = np.concatenate([b.t_col for b in blocks])
    vals = np.concatenate([b.t_val for b in blocks])
    return rows, cols, vals, m
