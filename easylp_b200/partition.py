"""Row-block partition of a large LP over the GPUs of one box (BASELINE.json north_star:
"Large LPs are partitioned by row blocks of A across the 8 GPUs of one box").

Rank g owns the contiguous rows [cuts[g], cuts[g+1]) of A (balanced by non-zeros, not by rows), the matching slices
of `sense`/`rhs`/y, and a replica of x, c, lb, ub.  Column ids stay global.  Pure host index arithmetic (numpy)."""
from __future__ import annotations

import numpy as np


def row_cuts(row_ptr, nranks):
    """cuts[0..nranks]: row boundaries that split the non-zeros as evenly as contiguous rows allow."""
    rp = np.asarray(row_ptr, dtype=np.int64)
    m = rp.size - 1
    nnz = int(rp[m]) if m >= 0 else 0
    cuts = [0]
    for g in range(1, nranks):
        cuts.append(int(np.searchsorted(rp, nnz * g / nranks, side="left")))
    cuts.append(m)
    for g in range(1, nranks + 1):            # monotone, inside [0, m]
        cuts[g] = min(max(cuts[g], cuts[g - 1]), m)
    return cuts


def row_block(p, rank, nranks):
    """The LP dict of rank `rank` (keys of oracle/gen.py): its rows, global columns.  Returns (block, r0, r1)."""
    rp = np.asarray(p["row_ptr"], dtype=np.int64)
    cuts = row_cuts(rp, nranks)
    r0, r1 = cuts[rank], cuts[rank + 1]
    q = dict(p)
    q["m"] = r1 - r0
    q["row_ptr"] = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
    q["col_idx"] = p["col_idx"][rp[r0]:rp[r1]]
    q["vals"] = p["vals"][rp[r0]:rp[r1]]
    q["sense"] = p["sense"][r0:r1]
    q["rhs"] = p["rhs"][r0:r1]
    return q, r0, r1


def lp_ranges(B, nranks):
    """Contiguous LP ranges of the batched path (no collective): [(b0, b1)] per rank."""
    base, extra = divmod(int(B), nranks)
    out, b = [], 0
    for g in range(nranks):
        e = b + base + (1 if g < extra else 0)
        out.append((b, e))
        b = e
    return out
