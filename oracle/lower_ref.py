"""CPU restatement of the lowered assembly (elp_expand_terms / elp_assemble_lowered).  TEST INFRASTRUCTURE ONLY.

Plain Python loops over the family descriptors of easylp_b200/lower.py::pack — the checker for the device expansion
kernel and for the two-level ordered fold.  What the stream must equal is the reference's own evaluation order:
`for` atoms in sequence (R/utils.R:50-53), `sum_for` grid rows with the first name fastest (R/utils.R:402-408), each
`sum_for` folded on its own before the atom's results are added or scaled (R/methods.R:82-111, 244-257).
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def expand(packed):
    """families -> the term stream (row, col, val, group), one Python iteration per term"""
    fam, n_fam, itab, dtab, _grp, _ng, n_low = packed
    row = np.zeros(n_low, np.int32)
    col = np.zeros(n_low, np.int32)
    val = np.zeros(n_low)
    grp = np.zeros(n_low, np.int32)
    for i in range(n_fam):
        f = fam[i]
        nl = f.n_loops
        ext = [f.extent[l] for l in range(nl)]
        for idx in range(f.count):
            rem, r, c, ci = idx, f.row0, f.col0, f.coef_tab
            for l in range(nl - 1, -1, -1):
                p = rem % ext[l]
                rem //= ext[l]
                r += int(itab[f.row_tab[l] + p]) if f.row_tab[l] >= 0 else f.row_stride[l] * p
                if f.col_tab[l] >= 0:
                    c += int(itab[f.col_tab[l] + p])
                ci += f.coef_stride[l] * p
            pos = f.out_offset + idx * f.out_stride
            row[pos], col[pos], val[pos], grp[pos] = r, c, dtab[ci], f.group
    return row, col, val, grp


def fold(row, col, val, grp, packed, m):
    """two-level ordered fold of a (group-major per entry) stream -> canonical CSR"""
    _fam, _nf, _itab, dtab, groups, _ng, _n = packed
    acc = {}
    order = {}
    for r, c, v, g in zip(row.tolist(), col.tolist(), val.tolist(), grp.tolist()):
        cell = acc.setdefault((r, c), [])
        if cell and cell[-1][0] == g:
            cell[-1][1] = cell[-1][1] + v
        else:
            assert not cell or cell[-1][0] < g, "stream is not group-major inside an entry"
            cell.append([g, v])
    rows = [[] for _ in range(m)]
    for (r, c), parts in acc.items():
        tot, have = 0.0, False
        for g, s in parts:
            if g != 0:
                G = groups[g]
                for k in range(G.n_mul):
                    s = s * float(dtab[G.mul_tab[k] + ((r - G.row0) if G.mul_per_row[k] else 0)])
            if s != 0.0:
                tot = tot + s if have else s
                have = True
        if have and tot != 0.0:
            rows[r].append((c, tot))
    rp, ci, vv = [0], [], []
    for r in range(m):
        for c, v in sorted(rows[r]):
            ci.append(c)
            vv.append(v)
        rp.append(len(ci))
    return np.asarray(rp, np.int32), np.asarray(ci, np.int32), np.asarray(vv, float)
