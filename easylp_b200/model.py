"""Host side of the EasyLP B200 solve path: the reference's R6 API on sparse term lists.

This module mirrors, name for name, the modelling interface of benet1one/EasyLP
(`easylp$new()/$var()/$con()/$min()/$max()/$solve()`, `sum_for`, `parameter`, the shadowed
`rowSums/colSums/rowMeans/colMeans/diag/apply`, `$alias/$uncon/$associate/$check_feasible/...`) so that
tests read like the reference's own (tests/testthat/*.R).  The R twin of this file is rpkg/R/*.R, which
binds the same C ABI through `.Call` (rpkg/src/r_glue.c); R is not installed in this image, so this
Python host is what the test-suite drives.  Citations are to /root/reference/.

What changed underneath: the reference keeps every `lp_var` as a DENSE `rows x n_var`
matrix (R/class.R:119-121) and `rbind`s dense rows into `constraint$mat` (R/utils.R:100).  Here an
`lp_var` is a list of (row, col, val) terms; three things happen on the GPU through include/easylp_abi.h:

  * `$con()` / `$min()` / `$max()`  -> elp_assemble_csr : stable radix sort of the (row, col) keys, ordered
    left-to-right fold of duplicates (exactly the `Reduce("+")` of `sum(...)`/`sum_for`, R/methods.R:244-257),
    zero drop, prefix-sum, scatter  ->  the canonical CSR of `constraint$mat` / `objective_fun`;
  * `$solve()`  -> elp_solve_lp : batched-simplex kernel when the dense tableau fits one SM's shared memory,
    PDLP otherwise (a size rule) — replacing lpSolveAPI (R/class.R:260-278);
  * `$check_feasible()` -> elp_check_feasible : `mat %*% sol` + compare_tol (R/class.R:533-540).

There is NO CPU solve path: without the CUDA library these calls raise.  Host arithmetic that stays here is
the per-operation algebra of the expression tree, kept bit-identical to the reference's dense arithmetic:
`x / k` multiplies by `1/k` (R/methods.R:163); unary minus and `k - x` negate `coef` only
(R/methods.R:155-159,192-194); `sum(x)` of a multi-row variable is R's colSums, a sequential long-double
accumulation; `[` keeps rows by membership in original order (R/methods.R:65).
Python cannot tell `2 >= x` from `x <= 2` (it calls x.__le__(2)); use compare(2, ">=", x) to get the
reference's row `-x >= -2` (vignettes/constraints.Rmd:225-230).
"""
from __future__ import annotations

import copy
import itertools
import sys
import warnings

import numpy as np

from . import _lib
from . import lower as _lower

LD = np.longdouble
# `for_` / `sum_for` first try ONE symbolic evaluation of their body (lower.py) and keep the rows as
# index-set descriptors that the device expands; bodies that cannot be traced run the reference's per-atom loop.
LOWERING = True
_I = np.int64


class EasyLpError(Exception):
    """stop() in the reference."""


def _identity(v):
    return v


def message(text, sink=None):
    """R's message(): to stderr, and recorded on the model so tests can expect it."""
    print(text, file=sys.stderr)
    if sink is not None:
        sink.append(text)


def _is_var(x):
    return isinstance(x, lp_var)


def _chr(v):
    if isinstance(v, (float, np.floating)) and float(v).is_integer():
        return str(int(v))
    return str(v)


def _as_vec(k):
    """R numeric vector in storage (column-major) order."""
    if isinstance(k, Param):
        return k.a.flatten(order="F").astype(float)
    a = np.asarray(k, dtype=float)
    return a.flatten(order="F") if a.ndim else a.reshape(1)


def _rsum(v):
    """R sum() of doubles: sequential long double accumulation."""
    v = np.asarray(v, dtype=float).ravel()
    return float(np.cumsum(v.astype(LD))[-1]) if v.size else 0.0


def _recycle(a, b):
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    if a.size == b.size:
        return a, b
    if a.size == 1:
        return np.repeat(a, b.size), b
    if b.size == 1:
        return a, np.repeat(b, a.size)
    raise EasyLpError("longer object length is not a multiple of shorter object length")


def _fold(row, col, val, dtype=float):
    """Left-to-right fold of duplicate (row, col) terms in emission order; returns canonical terms
    (sorted by row, then col; unique; exact zeros dropped).  dtype=LD gives R's colSums accumulation."""
    if row.size == 0:
        return row.astype(_I), col.astype(_I), val.astype(float)
    width = int(col.max()) + 1
    key = row.astype(_I) * width + col.astype(_I)
    order = np.argsort(key, kind="stable")
    k = key[order]
    v = val[order].astype(dtype)
    first = np.flatnonzero(np.r_[True, k[1:] != k[:-1]])
    counts = np.diff(np.r_[first, k.size])
    acc = v[first].copy()
    for t in range(1, int(counts.max())):
        sel = np.flatnonzero(counts > t)
        acc[sel] = acc[sel] + v[first[sel] + t]
    acc = acc.astype(float)
    kk = k[first]
    keep = acc != 0.0
    kk, acc = kk[keep], acc[keep]
    return kk // width, kk % width, acc


# ------------------------------------------------------------------------------------------------
class Param:
    """parameter(): named array (R/utils.R:356-375)."""

    def __init__(self, a, dimnames):
        self.a = np.asarray(a, dtype=float)
        self.dimnames = dimnames

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def __len__(self):
        return self.a.size

    def __repr__(self):
        return f"Param({self.a!r}, dimnames={self.dimnames})"

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if _lower.has_symbolic(key):          # inside a for/sum_for trace (lower.py)
            return _lower.param_getitem(self, key)
        pos = _positions(self.a.shape, self.dimnames, None, key)
        if len(key) == 1 and self.a.ndim > 1:
            out = self.a.flatten(order="F")[pos[0]]
        else:
            out = self.a[np.ix_(*pos)]
        out = np.asarray(out)
        return float(out.ravel()[0]) if out.size == 1 else out.squeeze()


def index_sets(mapping):
    """A list of index sets keyed by the values of another index — R: a named list used as `out[[v]]`, e.g. the arcs
    leaving node v.  Indexing with a value gives the plain list; inside a for/sum_for trace `sets[v]` stands for all of
    them, so `for (v in V) sum_for(a = out[[v]], ...)` lowers to one family over the (v, a) pairs (lower.py)."""
    return _lower.IndexSets(mapping)


def parameter(x, *sets, byrow=False, **named):
    sets = list(sets) + list(named.values())
    if not sets:
        raise EasyLpError("Parameter does not have any sets.")
    dims = tuple(len(s) for s in sets)
    x = np.asarray(x, dtype=float).ravel()
    if x.size == 1:
        x = np.repeat(x, int(np.prod(dims)))
    elif x.size != int(np.prod(dims)):
        raise EasyLpError("Dimensions of the parameter don't match dimensions of the sets.")
    dn = [[_chr(v) for v in s] for s in sets]
    if byrow:
        if len(sets) != 2:
            raise EasyLpError("Use 'byrow = TRUE' only with 2-dimensional arrays.")
        return Param(x.reshape(dims, order="C"), dn)
    return Param(x.reshape(dims, order="F"), dn)


def _positions(shape, dimnames, titles, key):
    """find_incorrect_index + R's `[` (R/utils.R:108-145): 0-based positions per subscript."""
    def one(ind, length, names):
        if ind is None or (isinstance(ind, slice) and ind == slice(None)):
            return np.arange(length)
        if isinstance(ind, (str, np.str_)):
            ind = [ind]
        if isinstance(ind, range):
            ind = list(ind)
        arr = np.asarray(ind)
        if arr.dtype == bool:
            return None
        if arr.dtype.kind in "iuf":
            arr = arr.ravel()
            if not (np.all(arr >= 1) and np.all(arr < length + 1)):
                return None
            return arr.astype(_I) - 1
        if arr.dtype.kind in "US":
            if names is None:
                return None
            out = []
            for s in arr.ravel().tolist():
                if s not in names:
                    return None
                out.append(names.index(s))
            return np.asarray(out, dtype=_I)
        return None

    if len(key) == 1:
        n = int(np.prod(shape))
        names = dimnames[0] if (dimnames is not None and len(shape) == 1) else None
        p = one(key[0], n, names)
        if p is None:
            raise EasyLpError("Invalid subscript")
        return [p]
    if len(key) != len(shape):
        raise EasyLpError("Invalid subscript: incorrect number of dimensions")
    out = []
    for d, k in enumerate(key):
        p = one(k, shape[d], dimnames[d] if dimnames is not None else None)
        if p is None:
            t = titles[d] if titles and titles[d] else d + 1
            raise EasyLpError(f"Invalid subscript on dimension '{t}'")
        out.append(p)
    return out


# ------------------------------------------------------------------------------------------------
class lp_con:
    """`lp_con` (R/methods.R:219-224): rows of terms, `dir`, `rhs`; names are attached by `$con`."""

    def __init__(self, nrow, t_row, t_col, t_val, dir, rhs):
        self.nrow, self.t_row, self.t_col, self.t_val = nrow, t_row, t_col, t_val
        self.dir, self.rhs = list(dir), np.asarray(rhs, dtype=float)
        self.names, self.rownames = [], []


class ForSplit(list):
    def __init__(self, items, variable, sequence):
        super().__init__(items)
        self.variable, self.sequence = variable, list(sequence)


def for_(body, **index):
    """`for (v in seq) body` inside `$con()` (R/utils.R:33-64).  Several indices nest, first outermost;
    write nested for_ calls when an inner range depends on an outer index (test-investments.R:35-37)."""
    if LOWERING and index:
        low = _lower.try_for(body, index)     # one symbolic evaluation instead of one per atom; None = not lowerable
        if low is not None:
            return low
    (var, seq), rest = next(iter(index.items())), dict(list(index.items())[1:])
    seq = list(seq)
    if rest:
        return ForSplit([for_(lambda _v=v, **kw: body(**{var: _v}, **kw), **rest) for v in seq], var, seq)
    return ForSplit([body(**{var: v}) for v in seq], var, seq)


class lp_var:
    """An affine vector expression `coef . x + add` as a term list (reference: R/class.R:161-174).

    t_row/t_col/t_val hold the non-zero coefficients.  `canonical` means sorted by (row, col) and unique;
    otherwise the list is a PENDING left-to-right sum (the un-evaluated `Reduce("+")` of `sum(a, b, ...)`),
    which the device assembly folds when the expression goes straight into `$con/$min/$max`, and which
    `_canon()` folds on the host if more algebra follows."""
    __array_ufunc__ = None

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __len__(self):                      # length.lp_var  R/methods.R:42
        return self.nrow

    @property
    def dim(self):
        return self.ind.shape if self.has_dim else None

    def copy(self):
        return copy.copy(self)

    # ---- term-list plumbing ----------------------------------------------------------------------
    def _canon(self):
        if not self.canonical:
            self.t_row, self.t_col, self.t_val = _fold(self.t_row, self.t_col, self.t_val)
            self.canonical = True
            self._ptr = None
        return self

    def _row_ptr(self):
        self._canon()
        if getattr(self, "_ptr", None) is None:
            self._ptr = np.searchsorted(self.t_row, np.arange(self.nrow + 1)).astype(_I)
        return self._ptr

    def _take_rows(self, rows):
        """terms of the given rows (ascending positions), renumbered 0..len(rows)-1"""
        ptr = self._row_ptr()
        if rows.size == 1:
            a, b = int(ptr[rows[0]]), int(ptr[rows[0] + 1])
            return np.zeros(b - a, _I), self.t_col[a:b], self.t_val[a:b]
        lens = ptr[rows + 1] - ptr[rows]
        total = int(lens.sum())
        new_row = np.repeat(np.arange(rows.size, dtype=_I), lens)
        if total == 0:
            return new_row, np.zeros(0, _I), np.zeros(0)
        offs = np.repeat(ptr[rows] - np.r_[0, np.cumsum(lens)[:-1]], lens)
        src = np.arange(total, dtype=_I) + offs
        return new_row, self.t_col[src], self.t_val[src]

    def dense(self, n):
        """dense `coef` (tests / printing only)"""
        x = self.copy()._canon()
        out = np.zeros((x.nrow, n))
        out[x.t_row, x.t_col] = x.t_val
        return out

    # ---- `[.lp_var`  R/methods.R:48-69 -----------------------------------------------------------
    def __getitem__(self, key):
        if not self.indexable:
            raise EasyLpError("Cannot index this result.")
        if not isinstance(key, tuple):
            key = (key,)
        if _lower.has_symbolic(key):          # inside a for/sum_for trace (lower.py)
            return _lower.var_getitem(self, key)
        pos = _positions(self.ind.shape, self.dimnames, self.dimtitles, key)
        x = self.copy()
        if len(key) == 1 and self.ind.ndim > 1:
            lin = pos[0]
            x.ind = self.ind.flatten(order="F")[lin]
            x.dimnames, x.dimtitles, x.has_dim = None, None, False
        else:
            x.ind = self.ind[np.ix_(*pos)]
            x.dimnames = ([[self.dimnames[d][i] for i in p] for d, p in enumerate(pos)]
                          if self.dimnames is not None else None)
            lin = np.ravel_multi_index(np.ix_(*pos), self.ind.shape, order="F").ravel() if len(pos) > 1 else pos[0]
        if self.ind.size != self.nrow:
            raise EasyLpError("Variable was wrongly indexed.")
        # rows <- is.element(old_ind, x$ind): membership, original order.  ids inside one `ind` are unique,
        # so the member rows are exactly the (sorted, de-duplicated) linear positions that were picked.
        rows = np.unique(np.asarray(lin, dtype=_I))
        x.raw = False
        x.t_row, x.t_col, x.t_val = self._take_rows(rows)
        x.canonical, x._ptr = True, None
        x.nrow = int(rows.size)
        x.add = self.add[rows]
        return x

    # ---- Ops  R/methods.R:114-199 ----------------------------------------------------------------
    def _checked(self):
        if np.isnan(self.t_val).any() or np.isnan(self.add).any():
            raise EasyLpError("Operation resulted in NA values")
        self.raw = False
        return self

    def __pos__(self):
        return self.copy()._checked()

    def __neg__(self):
        x = self.copy()
        x.t_val = -x.t_val                  # `add` is NOT negated (R/methods.R:155-159)
        return x._checked()

    def _scaled(self, kv, add_op):
        x = self.copy()._canon()
        if not np.all(np.isfinite(kv)):
            raise EasyLpError("Operation resulted in NA values")     # 0 * Inf in the dense reference
        nrow = x.nrow
        if nrow == 1 and kv.size > 1:       # horizontal_multiply: a 1-row variable is recycled (R/methods.R:84-85)
            nt = x.t_col.size
            x.t_row = np.repeat(np.arange(kv.size, dtype=_I), nt)
            x.t_col = np.tile(x.t_col, kv.size)
            x.t_val = np.tile(x.t_val, kv.size)
            nrow = kv.size
        elif kv.size == 1:
            kv = np.repeat(kv, nrow)
        if nrow != kv.size:
            raise EasyLpError("Linear variable must have the same length as multiplier Only values of size one are recycled.")
        v = x.t_val * kv[x.t_row]
        keep = v != 0.0
        x.t_row, x.t_col, x.t_val = x.t_row[keep], x.t_col[keep], v[keep]
        x.nrow, x._ptr = nrow, None
        x.add = add_op(*_recycle(x.add, kv))
        return x._checked()

    def __mul__(self, k):
        if _is_var(k):
            raise EasyLpError("Can't multiply or divide variables in a linear problem")
        kv = _as_vec(k)
        return self._scaled(kv, np.multiply)

    __rmul__ = __mul__

    def __truediv__(self, k):
        if _is_var(k):
            raise EasyLpError("Can't multiply or divide variables in a linear problem")
        kv = _as_vec(k)
        x = self._scaled(1.0 / kv, lambda a, b: a)          # coef * (1/k)  (R/methods.R:163)
        a, b = _recycle(self.add, kv)
        x.add = a / b                                       # add / k       (R/methods.R:164)
        return x._checked()

    def __rtruediv__(self, k):
        raise EasyLpError("Can't divide by a variable in a linear problem")

    def _plus_var(self, o, sign):
        x = self.copy()._canon()
        o = o.copy()._canon()
        xr, xc, xv, xn = x.t_row, x.t_col, x.t_val, x.nrow
        orow, oc, ov, on = o.t_row, o.t_col, sign * o.t_val, o.nrow
        if xn == 1 and on != 1:             # horizontal_mat_sum recycling (R/methods.R:100-103)
            xr, xc, xv, xn = np.repeat(np.arange(on, dtype=_I), xc.size), np.tile(xc, on), np.tile(xv, on), on
        if on == 1 and xn != 1:
            orow, oc, ov, on = np.repeat(np.arange(xn, dtype=_I), oc.size), np.tile(oc, xn), np.tile(ov, xn), xn
        if xn != on:
            raise EasyLpError("Linear variables must have the same length. Only values of size one are recycled.")
        x.t_row, x.t_col, x.t_val = _fold(np.r_[xr, orow], np.r_[xc, oc], np.r_[xv, ov])
        x.nrow, x._ptr = xn, None
        a, b = _recycle(x.add, o.add)
        x.add = a + b if sign > 0 else a - b
        return x._checked()

    def __add__(self, k):
        if _is_var(k):
            return self._plus_var(k, +1.0)
        x = self.copy()
        x.add = np.add(*_recycle(x.add, _as_vec(k)))
        return x._checked()

    __radd__ = __add__

    def __sub__(self, k):
        if _is_var(k):
            return self._plus_var(k, -1.0)
        x = self.copy()
        a, b = _recycle(x.add, _as_vec(k))
        x.add = a - b
        return x._checked()

    def __rsub__(self, k):                  # k - x: coef negated, then e2 + e1 (R/methods.R:192-194)
        x = self.copy()
        x.t_val = -x.t_val
        return x + k

    def __pow__(self, k):
        raise EasyLpError("Can't use operations '^', '%%', '%/%' in a linear problem")

    __mod__ = __floordiv__ = __rpow__ = __pow__

    def __invert__(self):                   # `!x`  (R/methods.R:140-147)
        if not self.binary:
            raise EasyLpError("Logical negation '!' only supported on binary variables.")
        return -self + 1

    def __abs__(self):
        raise EasyLpError("Function 'abs' is not linear. See how to use absolute values in linear programming here:\n"
                          "https://optimization.cbe.cornell.edu/index.php?title=Optimization_with_absolute_values")

    def __iter__(self):
        raise TypeError("lp_var is not iterable; use Sum()/sum_for() from easylp_b200.model")

    # ---- Compare  R/methods.R:200-225 ------------------------------------------------------------
    def __le__(self, o): return compare(self, "<=", o)
    def __ge__(self, o): return compare(self, ">=", o)
    def __lt__(self, o): return compare(self, "<", o)
    def __gt__(self, o): return compare(self, ">", o)
    def __eq__(self, o): return compare(self, "==", o)
    def __ne__(self, o): raise EasyLpError("Inequality '!=' not allowed in linear problems")
    __hash__ = None


def compare(e1, op, e2):
    """Compare_lp_var (R/methods.R:200-225), with the reference's operand order."""
    if op == "!=":
        raise EasyLpError("Inequality '!=' not allowed in linear problems")
    if op not in ("<=", ">=", "<", ">", "=="):
        raise EasyLpError(f"unknown comparison '{op}'")
    if not _is_var(e1) and not _is_var(e2):
        raise EasyLpError("not a linear comparison")
    if _is_var(e2):
        x = e1 - e2
        rhs = np.zeros(1)
    else:
        x = e1
        rhs = _as_vec(e2)
    if rhs.size == 1:
        rhs = np.repeat(rhs, x.nrow)
    if rhs.size != x.nrow:
        raise EasyLpError("length(rhs) == nrow(x$coef) is not TRUE")
    rhs = rhs - x.add
    return lp_con(x.nrow, x.t_row, x.t_col, x.t_val, [op] * x.nrow, rhs)


def Sum(x, *dots):
    """sum.lp_var (R/methods.R:244-257): each argument is column-summed on its own, then `Reduce("+")`.
    The Reduce is left PENDING as a term list in argument order — the device assembly folds it."""
    if _lower.is_traced(x, *dots):            # inside a for/sum_for trace (lower.py)
        return _lower.sym_sum(x, *dots)
    if not _is_var(x):
        if any(_is_var(d) for d in dots):
            raise EasyLpError("invalid 'type' (list) of argument")
        return _rsum(np.concatenate([_as_vec(x)] + [_as_vec(d) for d in dots]))
    if dots:
        first = Sum(x)
        rows, cols, vals = [first.t_row], [first.t_col], [first.t_val]
        add = first.add.copy()
        for d in dots:
            if _is_var(d):
                s = Sum(d)
                rows.append(s.t_row); cols.append(s.t_col); vals.append(s.t_val)
                add = add + s.add
            else:
                add = add + _rsum(_as_vec(d))
        out = first.copy()
        out.t_row, out.t_col, out.t_val = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
        out.canonical, out._ptr = len(rows) == 1, None
        out.add = add
        return out._checked()
    x = x.copy()._canon()
    if x.nrow > 1 and x.t_col.size:
        # colSums: per column, rows in order, long double accumulation, rounded once
        r, c, v = _fold(np.zeros(x.t_col.size, _I), x.t_col, x.t_val, dtype=LD) if _has_dup(x.t_col) else \
            (np.zeros(x.t_col.size, _I), *_sorted_by_col(x.t_col, x.t_val))
        x.t_row, x.t_col, x.t_val = r, c, v
    else:
        x.t_row = np.zeros(x.t_col.size, _I)
    x.nrow, x._ptr, x.canonical = 1, None, True
    x.add = np.array([_rsum(x.add)])
    x.indexable = False
    x.raw = False
    return x


def _has_dup(col):
    return np.unique(col).size != col.size


def _sorted_by_col(col, val):
    o = np.argsort(col, kind="stable")
    return col[o], val[o]


def mean(x):
    if _lower.is_traced(x):
        return Sum(x) / len(x)
    if not _is_var(x):
        return float(np.mean(_as_vec(x)))
    return Sum(x) / len(x)


def weighted_mean(x, w):
    if len(_as_vec(w)) != len(x):
        raise EasyLpError("'x' and 'w' must have the same length")
    return Sum(x * w) / Sum(w)


def cumsum(x):
    """Math.lp_var (R/methods.R:228-242): long-double running sum of `add`, double row recurrence of `coef`."""
    if not _is_var(x):
        return np.cumsum(_as_vec(x).astype(LD)).astype(float)
    x = x.copy()._canon()
    x.add = np.cumsum(x.add.astype(LD)).astype(float)
    cols = np.unique(x.t_col)
    if cols.size and x.nrow >= 2:
        sub = np.zeros((x.nrow, cols.size))
        sub[x.t_row, np.searchsorted(cols, x.t_col)] = x.t_val
        for i in range(1, x.nrow):
            sub[i, :] = sub[i, :] + sub[i - 1, :]
        r, c = np.nonzero(sub != 0.0)
        x.t_row, x.t_col, x.t_val = r.astype(_I), cols[c], sub[r, c]
        x._ptr = None
    x.raw = False
    return x


def sum_for(body, **index):
    """sum_for (R/utils.R:391-411): expand.grid with the first index fastest, evaluate, `do.call(sum, ...)`."""
    if not index:
        raise EasyLpError("No named indexing variables.")
    if LOWERING:
        low = _lower.try_sum_for(body, index)  # one symbolic evaluation instead of one per grid row
        if low is not None:
            return low
    names = list(index)
    seqs = [list(index[k]) for k in names]
    result = []
    for cell in itertools.product(*reversed(seqs)):
        kw = dict(zip(reversed(names), cell))
        result.append(body(**kw))
    return Sum(*result)


# ---- shadowed base functions (R/utils.R:236-333) -------------------------------------------------
def _ensure_not_con(x, fun):
    if isinstance(x, lp_con):
        raise EasyLpError(f"Cannot apply function '{fun}' to a constraint.\n"
                          f"Did you accidentally write the constraint inside '{fun}()'?")


def diag(x):
    _ensure_not_con(x, "diag")
    if not _is_var(x):
        return np.diag(np.asarray(x))
    y = x[np.diag(x.ind)]       # the reference subscripts with the ids themselves (R/utils.R:243)
    y.raw = False
    return y


def apply(X, MARGIN, FUN):
    _ensure_not_con(X, "apply")
    if not _is_var(X):
        raise EasyLpError("apply(): only linear variables are handled here; use numpy for plain arrays")
    if isinstance(MARGIN, str) or (isinstance(MARGIN, (list, tuple)) and MARGIN and isinstance(MARGIN[0], str)):
        raise EasyLpError("Not all elements of 'MARGIN' are names of dimensions.")
    MARGIN = [int(v) for v in np.atleast_1d(np.asarray(MARGIN))]
    nd = X.ind.ndim
    if any(v < 1 or v > nd for v in MARGIN):
        raise EasyLpError("'MARGIN' does not match dim(X).")
    mdims = [X.ind.shape[v - 1] for v in MARGIN]
    cells = list(itertools.product(*[range(1, d + 1) for d in reversed(mdims)]))
    rows, cols, vals = [], [], []
    add = np.zeros(len(cells))
    for k, cell in enumerate(cells):
        cell = list(reversed(cell))                     # first margin fastest (expand.grid)
        ind = [np.arange(1, d + 1) for d in X.ind.shape]
        for v, c in zip(MARGIN, cell):
            ind[v - 1] = c
        z = FUN(X[tuple(ind)])
        if not _is_var(z) or z.nrow != 1:
            raise EasyLpError("number of items to replace is not a multiple of replacement length")
        z._canon()
        rows.append(np.full(z.t_col.size, k, _I)); cols.append(z.t_col); vals.append(z.t_val)
        add[k] = z.add[0]
    out = X.copy()
    out.ind = np.arange(1, len(cells) + 1).reshape(mdims, order="F")
    out.dimnames = [X.dimnames[v - 1] for v in MARGIN] if X.dimnames is not None else None
    out.dimtitles = [X.dimtitles[v - 1] for v in MARGIN] if X.dimtitles is not None else None
    out.has_dim = True
    out.t_row = np.concatenate(rows) if rows else np.zeros(0, _I)
    out.t_col = np.concatenate(cols) if cols else np.zeros(0, _I)
    out.t_val = np.concatenate(vals) if vals else np.zeros(0)
    out.canonical, out._ptr, out.nrow = True, None, len(cells)
    out.add, out.raw = add, False
    return out


def rowSums(x):
    _ensure_not_con(x, "rowSums")
    return apply(x, 1, Sum) if _is_var(x) else np.asarray(x).sum(axis=1)


def colSums(x):
    _ensure_not_con(x, "colSums")
    return apply(x, 2, Sum) if _is_var(x) else np.asarray(x).sum(axis=0)


def rowMeans(x):
    _ensure_not_con(x, "rowMeans")
    return apply(x, 1, mean) if _is_var(x) else np.asarray(x).mean(axis=1)


def colMeans(x):
    _ensure_not_con(x, "colMeans")
    return apply(x, 2, mean) if _is_var(x) else np.asarray(x).mean(axis=0)


# ------------------------------------------------------------------------------------------------
def name_constraint(con, name):         # R/utils.R:154-165
    n = con.nrow
    if not name:
        con.names = [""] * n
        con.rownames = [""] * n
        return con
    con.names = [name] * n
    con.rownames = [f"{name}[{k}]" for k in range(1, n + 1)] if n > 1 else [name]
    return con


def flatten_for_split(split, init_name=""):     # R/utils.R:66-94
    atoms = []

    def add(x, name):
        if isinstance(x, ForSplit):
            name = name.replace("]", ",", 1)
            for k, item in enumerate(x):
                add(item, f"{name}{x.variable}={_chr(x.sequence[k])}]")
        elif isinstance(x, _lower.LoweredFor):         # an inner `for` that traced while the outer one ran per atom
            blk = copy.copy(x.block)
            blk.init_name, blk.head = init_name, name.replace("]", ",", 1)
            atoms.append((name, blk))
        else:
            if isinstance(x, lp_con):
                x = name_constraint(x, name)
                x.names = [init_name] * x.nrow
            atoms.append((name, x))

    add(split, (init_name or "") + "[")
    return atoms


def large_to_infinity(x, threshold=1e30):       # R/utils.R:172-176
    x = np.array(x, dtype=float, copy=True)
    x[x >= threshold] = np.inf
    x[x <= -threshold] = -np.inf
    return x


_SENSE = {"<=": _lib.LE, "<": _lib.LE, ">=": _lib.GE, ">": _lib.GE, "==": _lib.EQ}   # add.constraint: "<" ~ "<="
_STRICT = {"<=": 0, ">=": 1, "==": 2, "<": 3, ">": 4}                                   # compare_tol keeps strictness


class Constraint:
    """`easylp$constraint` (R/class.R:56-61).  `mat` is the canonical CSR built on the device."""

    def __init__(self, lp):
        self._lp = lp

    @property
    def mat(self):
        return self._lp._csr()

    @property
    def dir(self):
        return [d for b in self._lp._blocks for d in b.dir]

    @property
    def rhs(self):
        bl = self._lp._blocks
        return np.concatenate([b.rhs for b in bl]) if bl else np.zeros(0)

    @property
    def names(self):
        return [s for b in self._lp._blocks for s in b.names]

    @property
    def rownames(self):
        return [s for b in self._lp._blocks for s in b.rownames]

    def todense(self):
        rp, ci, v = self.mat
        out = np.zeros((rp.size - 1, self._lp.nvar))
        for i in range(rp.size - 1):
            out[i, ci[rp[i]:rp[i + 1]]] = v[rp[i]:rp[i + 1]]
        return out


class easylp:
    """R6 class `easylp` (R/class.R:51-648) over the B200 C ABI."""

    def __init__(self):
        self.variables = {}
        self.aliases = {}
        self.constraint = Constraint(self)
        self.objective_add = 0.0
        self.objective_transform = None
        self._obj_terms = (np.zeros(0, _I), np.zeros(0))     # pending terms of the objective row
        self._obj_cache = None
        self.pointer = None                  # reference: lpSolveAPI handle (R/class.R:66); here: last elp_stats
        self.messages = []
        self._blocks = []
        self._model = None                   # device-resident CSR (elp_model handle), rebuilt whenever the rows change
        self._cache_csr = None
        self._n_var = 0
        self._dir = "min"
        self._sol = np.zeros(0)
        self._objval = np.nan
        self._stat = "unsolved"

    @classmethod
    def new(cls):
        return cls()

    def clone(self):
        return copy.deepcopy(self)

    def __getstate__(self):                 # pickling = saveRDS(): the device copy is a rebuildable cache, never state
        state = dict(self.__dict__)
        state["_model"] = None
        return state

    # ---- $var  R/class.R:85-179 ------------------------------------------------------------------
    def var(self, name, *sets, integer=False, binary=False, lower=-np.inf, upper=np.inf, **named):
        if not isinstance(name, str):
            raise EasyLpError("is_scalar_character(name) is not TRUE")
        if name in self.variables:
            raise EasyLpError(f"Variable '{name}' already defined in this model.")
        if lower > upper:
            warnings.warn("Lower bound is higher than upper bound. Problem will be unfeasible.")
        if binary:
            integer = False
            if lower != -np.inf or upper != np.inf:
                warnings.warn(f"Ignoring bounds for binary variable {name}")
            lower, upper = 0.0, 1.0
        titles = [""] * len(sets) + list(named)
        sets = [list(s) for s in sets] + [list(s) for s in named.values()]
        if not sets:
            sets, titles = [[""]], ["scalar"]
        dims = tuple(len(s) for s in sets)
        ln = int(np.prod(dims))
        ids = np.arange(1, ln + 1, dtype=_I) + self._n_var
        x = lp_var(name=name, ind=ids.reshape(dims, order="F"), dimnames=[[_chr(v) for v in s] for s in sets],
                   dimtitles=titles, has_dim=True, type="integer" if integer else ("binary" if binary else "real"),
                   integer=bool(integer), binary=bool(binary), bound=[float(lower), float(upper)], indexable=True,
                   raw=True, nrow=ln, t_row=np.arange(ln, dtype=_I), t_col=ids - 1, t_val=np.ones(ln),
                   canonical=True, _ptr=None, add=np.zeros(ln))
        self._obj_cache = None
        self._sol = np.concatenate([self._sol, np.zeros(ln)])
        if lower > 0 or upper < 0:
            self.reset_solution()
        self.variables[name] = x
        self._n_var += ln
        self._cache = None
        return x

    def __getitem__(self, name):
        return self.aliases[name] if name in self.aliases else self.variables[name]

    # ---- $alias  R/class.R:362-368 ---------------------------------------------------------------
    def alias(self, *unnamed, **named):
        if unnamed:
            raise EasyLpError("Aliases must be named.")
        self.aliases.update(named)

    # ---- $con  R/class.R:189-220 -----------------------------------------------------------------
    def con(self, *unnamed, **named):
        items = [(None, c) for c in unnamed] + list(named.items())
        for k, (name, c) in enumerate(items, 1):
            ref = name or k
            if callable(c):
                try:
                    c = c()
                except Exception as e:
                    raise EasyLpError(f"Constraint '{ref}' evaluated to an error:\n{e}") from e
            if isinstance(c, _lower.LoweredFor):      # rows kept as index-set descriptors, expanded on the device
                blk = copy.copy(c.block)
                blk.init_name, blk.head = name or "", (name or "") + "["
                self._blocks.append(blk)
                continue
            if isinstance(c, ForSplit):
                split = flatten_for_split(c, name or "")
                if not split or not isinstance(split[0][1], (lp_con, _lower.LoweredCon)):
                    raise EasyLpError("Constraint did not evaluate to an (in)equality.")
                for _, atom in split:
                    if not isinstance(atom, (lp_con, _lower.LoweredCon)):
                        raise EasyLpError("is_lp_con(con) is not TRUE")
                    self._blocks.append(atom)
                continue
            if not isinstance(c, lp_con):
                raise EasyLpError(f"Constraint '{ref}' did not evaluate to an (in)equality.")
            if c.nrow == 0:
                warnings.warn(f"Constraint '{ref}' is empty.")
                continue
            self._blocks.append(name_constraint(c, name))
        self._cache = None
        self.check_feasible()
        return self

    def uncon(self, name):               # R/class.R:308-316
        if not isinstance(name, (str, list, tuple)):
            raise EasyLpError("Use the name <character> of a constraint to remove it.")
        names = [name] if isinstance(name, str) else list(name)
        self._blocks = [b for b in self._blocks if not (b.names and b.names[0] in names)]
        self._cache = None
        return self

    # ---- device assembly ---------------------------------------------------------------------------
    def _dir_codes(self, table):
        """int8 code of every row's direction; lowered blocks have one direction for all their rows"""
        parts = [np.full(b.nrow, table[b._dir], dtype=np.int8) if isinstance(b, _lower.LoweredCon)
                 else np.array([table[d] for d in b.dir], dtype=np.int8) for b in self._blocks]
        return np.concatenate(parts) if parts else np.zeros(0, np.int8)

    def _device_model(self):
        """the model's canonical CSR, assembled on the device and KEPT there (elp_model_assemble): eager blocks hand over
        their term lists, lowered blocks their index-set families; `$solve()` runs on this handle, `_csr()` copies it back"""
        if self._model is None:
            offs = np.cumsum([0] + [b.nrow for b in self._blocks])
            m = int(offs[-1])
            eager = [(b, o) for b, o in zip(self._blocks, offs) if not isinstance(b, _lower.LoweredCon)]
            lowered = [(b, int(o)) for b, o in zip(self._blocks, offs) if isinstance(b, _lower.LoweredCon)]
            rows = np.concatenate([b.t_row + o for b, o in eager]) if eager else np.zeros(0, _I)
            cols = np.concatenate([b.t_col for b, _ in eager]) if eager else np.zeros(0, _I)
            vals = np.concatenate([b.t_val for b, _ in eager]) if eager else np.zeros(0)
            self._model = _lib.Model(rows, cols, vals, _lower.pack(lowered), m, self._n_var)
            self.assembly_stats = self._model.stats
        return self._model

    @property
    def _cache(self):
        return self._cache_csr

    @_cache.setter
    def _cache(self, value):                 # every `self._cache = None` of the methods below also drops the device copy
        self._cache_csr = value
        if value is None and getattr(self, "_model", None) is not None:
            self._model.close()
            self._model = None

    def _csr(self):
        """canonical CSR of `constraint$mat` on the host: sort + ordered fold + scan + scatter ran on the device
        (elp_model_assemble); this copies the result back once"""
        if self._cache is None:
            m = sum(b.nrow for b in self._blocks)
            if m == 0:
                self._cache = (np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
            else:
                self._cache = self._device_model().csr()
        return self._cache

    # ---- $min / $max  R/class.R:230-246, 509-531 ---------------------------------------------------
    def _set_objective(self, x, transform):
        if isinstance(x, lp_con):
            raise EasyLpError("Objective function evaluated to a constraint. It must evaluate to a variable or sum of variables.")
        if not _is_var(x):
            raise EasyLpError("Objective function didn't evaluate to a variable or sum of variables.")
        if len(x) == 0:
            raise EasyLpError("Objective function doesn't contain any variables.")
        if len(x) > 1:
            raise EasyLpError("Objective function contains multiple variables. Please wrap them in a sum().")
        self._obj_terms = (x.t_col.astype(_I), x.t_val.astype(float))
        self._obj_cache = None
        self.objective_add = float(x.add[0])
        if transform is not None:           # identity cannot decrease; skip the 64-point probe (and the assembly)
            lo, up = self._objective_bounds(self.objective_fun, self.objective_add)
            _warn_decreasing_transformation(transform, lo, up)
        self.objective_transform = transform if transform is not None else _identity
        self.reset_solution()
        return self

    @property
    def objective_fun(self):
        """`objective_fun` (R/class.R:63): the dense cost vector; its terms are folded on the device"""
        if self._obj_cache is None:
            c = np.zeros(self._n_var)
            cols, vals = self._obj_terms
            if cols.size:
                _, ci, v, _ = _lib.assemble_csr(np.zeros(cols.size, np.int32), cols, vals, 1, self._n_var)
                c[ci] = v
            self._obj_cache = c
        return self._obj_cache

    def _objective_bounds(self, c, add):
        """update_bounds (R/utils.R:177-197) for the one-row objective"""
        lb, ub = self._bounds()
        nz = c != 0
        with np.errstate(invalid="ignore"):
            a, b = c[nz] * lb[nz], c[nz] * ub[nz]
        a = np.where(np.isnan(a), 0.0, a)
        b = np.where(np.isnan(b), 0.0, b)
        return float(np.minimum(a, b).sum() + add), float(np.maximum(a, b).sum() + add)

    def min(self, objective, transform=None):
        self._dir = "min"
        return self._set_objective(objective, transform)

    def max(self, objective, transform=None):
        self._dir = "max"
        return self._set_objective(objective, transform)

    def _bounds(self):
        if not self.variables:
            return np.zeros(0), np.zeros(0)
        lb = np.concatenate([np.repeat(v.bound[0], v.ind.size) for v in self.variables.values()])
        ub = np.concatenate([np.repeat(v.bound[1], v.ind.size) for v in self.variables.values()])
        return lb, ub

    # ---- $solve  R/class.R:251-302 -----------------------------------------------------------------
    def solve(self, **control):
        if self._n_var == 0:
            raise EasyLpError("Problem contains no variables.")
        if np.all(self.objective_fun == 0):
            raise EasyLpError("Must specify objective function.")
        if self._dir not in ("min", "max"):
            raise EasyLpError("Direction must be either 'min' or 'max'.")
        opt = _lib.default_options()
        for k, v in control.items():             # the `...` of lp.control() (R/class.R:262)
            if k == "timeout":
                opt.time_limit_s = float(v)
            elif k == "verbose":
                opt.verbose = int(v) if not isinstance(v, str) else int(v not in ("neutral", "critical", "severe"))
            elif k == "gpu_tol":
                opt.eps_rel = float(v)
            elif k == "epsilon":                 # lp_solve's integer-rounding tolerance: nothing to map it to
                warnings.warn("lp.control option 'epsilon' is lp_solve's integer-rounding tolerance and is ignored; "
                              "the GPU path's optimality tolerance is gpu_tol")
            elif k == "gpu_devices":             # spread one large solve over this many GPUs of the box
                opt.devices = int(v)
            elif k == "gpu_max_iter":
                opt.max_iter = int(v)
            elif k == "gpu_method":
                opt.method = {"auto": 0, "simplex": 1, "pdlp": 2}[v] if isinstance(v, str) else int(v)
            elif k == "gpu_transpose":           # PDLP: "gather" = bit-reproducible CSC SpMV, "scatter" = fp64 reductions
                opt.transpose = {"auto": 0, "gather": 1, "scatter": 2}[v] if isinstance(v, str) else int(v)
            else:
                warnings.warn(f"lp.control option '{k}' has no meaning on the GPU path and is ignored")
        m = sum(b.nrow for b in self._blocks)
        lb, ub = self._bounds()
        sense = self._dir_codes(_SENSE)
        if self.any_integer():
            # set.type(prob, columns, "integer" | "binary") + lp_solve's branch and bound (R/class.R:264-276): the tree's
            # frontiers go through the batched simplex kernel (csrc/mip.cu)
            is_int = np.concatenate([np.repeat(bool(v.integer or v.binary), v.ind.size) for v in self.variables.values()])
            rp, ci, v = self._csr()
            r = _lib.solve_mip(m, self._n_var, rp, ci, v, sense, self.constraint.rhs, self.objective_fun, lb, ub, is_int,
                               maximize=self._dir == "max", options=opt)
        elif m > 0:       # the matrix stays in HBM between `$con()` and `$solve()`
            r = self._device_model().solve(sense, self.constraint.rhs, self.objective_fun, lb, ub,
                                           maximize=self._dir == "max", options=opt)
        else:
            rp, ci, v = self._csr()
            r = _lib.solve_lp(m, self._n_var, rp, ci, v, sense, self.constraint.rhs, self.objective_fun, lb, ub,
                              maximize=self._dir == "max", options=opt)
        self._objval = float(large_to_infinity([r.objval])[0])
        self._sol = large_to_infinity(r.x)
        self._stat = r.status_string
        self.duals = r.y
        for x in self.variables.values():
            if x.bound[0] > x.bound[1]:
                self._stat = "unfeasible"
        self.pointer = r.stats
        return self

    # ---- $check_feasible  R/class.R:375-390, 533-540 -----------------------------------------------
    def check_feasible(self, tol=2e-8):
        if self._stat == "unsolved":
            return self
        rp, ci, v = self._csr()
        m = rp.size - 1
        if m <= 0:
            raise EasyLpError("nrow(mat) > 0L is not TRUE")
        strict = self._dir_codes(_STRICT)
        sol = np.where(np.isfinite(self._sol), self._sol, 0.0) if not np.all(np.isfinite(self._sol)) else self._sol
        feas = _lib.check_feasible(m, self._n_var, rp, ci, v, sol, strict, self.constraint.rhs, tol)
        if not feas.all():
            nam = [r if r else str(i + 1) for i, r in enumerate(self.constraint.rownames)]
            unfeas = ",".join(n for n, f in zip(nam, feas) if not f)
            message(f"Constrainsts: {unfeas}; are unfeasible. Use easylp$solve() to find a new solution.", self.messages)
            self.reset_solution()
        return self

    def check_solved(self):
        if self._stat == "unsolved":
            raise EasyLpError("Linear Problem has not been solved. Use easylp$solve().")

    def any_integer(self):
        return any(v.integer or v.binary for v in self.variables.values())

    def reset_solution(self):
        self._stat = "unsolved"
        self._sol = np.zeros(self._n_var)
        self._objval = np.nan
        return self

    # ---- $associate  R/class.R:332-358 -------------------------------------------------------------
    def associate(self, x, binary, max1=None, max0=None, min1=None, min0=None):
        lo, up = self._expr_bounds(x)
        max1 = up if max1 is None else max1
        max0 = lo if max0 is None else max0
        min1 = lo if min1 is None else min1
        min0 = lo if min0 is None else min0
        if not all(np.isfinite(v) for v in (max1, max0, min1, min0)):
            raise EasyLpError("is.finite(max1), is.finite(max0), is.finite(min1), is.finite(min0) are not all TRUE")
        if not binary.binary:
            warnings.warn("Variable is not binary.")
        if max1 != up or max0 != up:
            self.con(assoc_max=x <= max0 + (max1 - max0) * binary)
        if min1 != lo or min0 != lo:
            self.con(assoc_min=x >= min0 + (min1 - min0) * binary)
        return self

    def _expr_bounds(self, x):
        """update_bounds (R/utils.R:177-197) for a general expression"""
        x = x.copy()._canon()
        lb, ub = self._bounds()
        with np.errstate(invalid="ignore"):
            a, b = x.t_val * lb[x.t_col], x.t_val * ub[x.t_col]
        a = np.where(np.isnan(a), 0.0, a)
        b = np.where(np.isnan(b), 0.0, b)
        up = np.bincount(x.t_row, weights=np.maximum(a, b), minlength=x.nrow) + x.add
        lo = np.bincount(x.t_row, weights=np.minimum(a, b), minlength=x.nrow) + x.add
        return float(lo.min()), float(up.max())

    # ---- $test  R/class.R:435-467 (evaluation already happened in Python; returns the pieces) -------
    def test(self, *unnamed, **named):
        out = {}
        for k, (name, c) in enumerate([(None, c) for c in unnamed] + list(named.items()), 1):
            if callable(c):
                global LOWERING
                keep, LOWERING = LOWERING, False         # `$test` shows the evaluated atoms: no lowering here
                try:
                    c = c()
                except Exception as e:       # tryCatch(..., error = identity)
                    c = e
                finally:
                    LOWERING = keep
            if isinstance(c, ForSplit):
                c = [a for _, a in flatten_for_split(c, name or "")]
            elif isinstance(c, _lower.LoweredFor):       # evaluated by the caller with lowering on: the block of rows
                c = c.block
            out[name or k] = c
        return out

    def import_solution(self, envir, silent=False):
        self.check_solved()
        envir.update(self.solution)
        if not silent:
            message("Solution imported.", self.messages)
        return self

    # ---- active bindings  R/class.R:566-647 --------------------------------------------------------
    @property
    def nvar(self):
        return self._n_var

    @property
    def ncon(self):
        return sum(b.nrow for b in self._blocks)

    @property
    def direction(self):
        return self._dir

    @direction.setter
    def direction(self, arg):
        if isinstance(arg, str) and arg.lower() in ("min", "max"):
            self._dir = arg.lower()
        else:
            raise EasyLpError("Direction must be either 'min' or 'max'.")

    @property
    def solution(self):
        if self._stat != "optimal":
            warnings.warn("Problem is not optimal.\n")
        out = {}
        for name, x in self.variables.items():
            vals = self._sol[x.ind.flatten(order="F") - 1]
            out[name] = float(vals[0]) if x.ind.size == 1 else vals.reshape(x.ind.shape, order="F")
        return out

    @property
    def objective_value(self):
        self.check_solved()
        t = self.objective_transform or (lambda v: v)
        return t(self._objval + self.objective_add)

    @property
    def objective_value_raw(self):
        self.check_solved()
        return self._objval

    @property
    def status(self):
        return self._stat

    # ---- $sensitivity_objective / $sensitivity_rhs  R/class.R:613-646 ----------------------------------------
    def _sensitivity(self):
        if self._stat != "optimal":
            raise EasyLpError("Problem is not optimal.")
        if self.any_integer():
            raise EasyLpError("Sensitivity unavailable for problems with integer/binary variables")
        lb, ub = self._bounds()
        rp, ci, v = self._csr()
        m = sum(b.nrow for b in self._blocks)
        return _lib.sensitivity(m, self._n_var, rp, ci, v, self._dir_codes(_SENSE), self.constraint.rhs, self.objective_fun,
                                lb, ub, maximize=self._dir == "max")

    @property
    def sensitivity_objective(self):
        """array [variable, (Lower, Current, Upper)] like R/class.R:619-628: the range of every objective coefficient
        over which the optimal basis stays optimal (elp_sensitivity; infinite ends are +/-inf as after large_to_infinity)"""
        _, _, _, of, ot, _, _, _ = self._sensitivity()
        return np.column_stack([of, self.objective_fun, ot])

    @property
    def sensitivity_rhs(self):
        """array [constraint, (Lower, Current, Upper)] like R/class.R:636-645: the range of every right-hand side over
        which the optimal basis stays feasible"""
        _, _, _, _, _, rf, rt, _ = self._sensitivity()
        return np.column_stack([rf, self.constraint.rhs, rt])

    def variable_names(self):
        """name_variable (R/utils.R:147-153), for every column"""
        out = []
        for name, x in self.variables.items():
            if x.ind.ndim == 1 and x.ind.size == 1:
                out.append(name)
                continue
            grids = itertools.product(*reversed(x.dimnames))
            out += [f"{name}[{','.join(reversed(g))}]" for g in grids]
        return out

    # ---- lp_solve LP-format export ------------------------------------------------
    def write_lp(self, path=None):
        """The assembled model in lp_solve's LP format (17-digit coefficients, the canonical CSR row by row), so that a
        machine with R + lpSolveAPI can solve byte-identical input with `read.lp()` and time / check the reference
        backend at sizes the dense DSL cannot build.  Returns the text; writes it to `path` when given."""
        rp, ci, v = self._csr()
        names = [_lp_name(s) for s in self.variable_names()]
        c = self.objective_fun
        fmt = lambda a: repr(float(a))

        def lin(cols, vals):
            if len(cols) == 0:
                return "0"
            return " ".join(f"{'+' if a >= 0 else '-'}{fmt(abs(a))} {names[j]}" for j, a in zip(cols, vals))

        nz = np.flatnonzero(c)
        out = [f"/* easylp_b200 export: {self._n_var} variables, {rp.size - 1} constraints */",
               f"{'max' if self._dir == 'max' else 'min'}: {lin(nz, c[nz])};", ""]
        ops = {"<=": "<=", "<": "<=", ">=": ">=", ">": ">=", "==": "="}
        rown = self.constraint.rownames
        for i in range(rp.size - 1):
            a, b = rp[i], rp[i + 1]
            label = f"{_lp_name(rown[i])}: " if rown[i] else ""
            out.append(f"{label}{lin(ci[a:b], v[a:b])} {ops[self.constraint.dir[i]]} {fmt(self.constraint.rhs[i])};")
        out.append("")
        lb, ub = self._bounds()
        free = [names[j] for j in range(self._n_var) if lb[j] == -np.inf and ub[j] == np.inf]
        for j in range(self._n_var):
            if lb[j] == -np.inf and ub[j] == np.inf:
                continue
            if lb[j] == -np.inf:
                out.append(f"{names[j]} >= -1e30;")
                out.append(f"{names[j]} <= {fmt(ub[j])};")
            elif ub[j] == np.inf:
                if lb[j] != 0.0:
                    out.append(f"{names[j]} >= {fmt(lb[j])};")
            else:
                out.append(f"{fmt(lb[j])} <= {names[j]} <= {fmt(ub[j])};")
        if free:
            out.append("free " + ", ".join(free) + ";")
        ints = [n_ for n_, x in zip(names, (t for x in self.variables.values() for t in [x] * x.ind.size)) if x.integer or x.binary]
        if ints:
            out.append("int " + ", ".join(ints) + ";")
        text = "\n".join(out) + "\n"
        if path is not None:
            with open(path, "w") as fh:
                fh.write(text)
        return text

    def __repr__(self):                 # $print  R/class.R:470-494
        s = f"Easy Linear Problem \nStatus: {self._stat}"
        if self._stat != "optimal":
            return s
        s += f"\nObjective Value = {self._objval}"
        if self.objective_add != 0:
            s += f" {'+' if self.objective_add > 0 else '-'} {abs(self.objective_add)} = {self.objective_value}"
        return s + f"\n\nSolution:\n\n{self.solution}"


def _lp_name(s):
    """an identifier lp_solve's LP parser accepts: letters, digits and _ [ ] . only"""
    return "".join(ch if (ch.isascii() and (ch.isalnum() or ch in "_[].")) else "_" for ch in s)


def _warn_decreasing_transformation(f, lower, upper):    # R/utils.R:199-217
    lower = lower if np.isfinite(lower) else -1e3
    upper = upper if np.isfinite(upper) else max(1e3, lower + 2e3)
    last = -np.inf
    for x in np.linspace(lower, upper, 64):
        try:
            with np.errstate(all="ignore"):
                y = f(x)
            if y is None or np.isnan(y):
                raise ValueError
        except Exception:
            warnings.warn("Could not ensure transformation is increasing within bounds of objective value.")
            return
        if y < last:
            warnings.warn("Transformation decreases within bounds of objective value."
                          "Solution might not be optimal with linear methods.")
            return
        last = y


def solve_batch(A, b, c, lower=0.0, upper=np.inf, dir="<=", sense="min", **control):
    """Additive batch entry point (BASELINE config 3): many independent small dense LPs
    `min/max c_k'x  s.t.  A_k x dir b_k, lower <= x <= upper`, one LP per warp (per CTA for the larger shapes).  Returns (status strings,
    objective values, solutions)."""
    A = np.asarray(A, dtype=float)
    B, m, n = A.shape
    lb = np.broadcast_to(np.asarray(lower, dtype=float), (B, n))
    ub = np.broadcast_to(np.asarray(upper, dtype=float), (B, n))
    if isinstance(dir, str):
        sn = np.full((B, m), _SENSE[dir], dtype=np.int8)
    else:
        sn = np.broadcast_to(np.vectorize(_SENSE.get)(np.asarray(dir)).astype(np.int8), (B, m))
    opt = _lib.default_options()
    if "gpu_max_iter" in control:
        opt.max_iter = int(control["gpu_max_iter"])
    status, obj, x, st = _lib.solve_batch(A, b, c, lb, ub, sn, maximize=sense == "max", options=opt)
    return [_lib.status_string(int(s)) for s in status], large_to_infinity(obj), large_to_infinity(x), st
