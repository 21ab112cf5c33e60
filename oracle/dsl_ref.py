"""Dense CPU restatement of EasyLP's modelling DSL — TEST INFRASTRUCTURE ONLY ("the assembly oracle").

Only tests/ may import this file.  Nothing under easylp_b200/ does.

It restates, statement for statement, the reference's *dense* R arithmetic that builds `constraint$mat`,
`dir`, `rhs`, `objective_fun`, `objective_add` and the bounds, so that the product's sparse host code +
device CSR assembly can be compared with it BIT-EXACTLY (canonical CSR = entries of the dense matrix
that are != 0, row-major, ascending column).  Every function cites the reference lines it follows
(paths relative to /root/reference/).  R idioms are kept where they decide bits:

  * arrays are column-major, ids are 1-based (`ind[] <- 1:length(ind) + n_var`, R/class.R:112-113);
  * `x / k` is `coef * (1/k)` (R/methods.R:163), scalars are applied eagerly, row by row;
  * `sum(x)` is R's colSums — a sequential LONG DOUBLE accumulation per column (src/main/array.c) —
    and `sum(a, b, ...)` is `Reduce("+")` over the individually summed pieces (R/methods.R:244-257);
  * `cumsum` on `add` is R's long-double running sum, on `coef` a double row recurrence (R/methods.R:228-242);
  * unary minus and `k - x` negate `coef` only, never `add` (R/methods.R:155-159, 192-194);
  * `[` keeps rows by MEMBERSHIP in original order (R/methods.R:48-69).

The Python surface (easylp, parameter, sum_for, for_, Sum, mean, ...) is the same as the product's
easylp_b200.model so one model-building function can be run against both.  Python cannot see that
`2 >= x` was written with the number on the left (it calls x.__le__(2)); use compare(2, ">=", x) for
the reference's exact row in that case.

Parity pinning: goldens G1-G7 of SURVEY.md §8c are checked against this file in tests/test_dsl_oracle.py
(README row values, test-DOP objective 3 985 000 via the simplex oracle, row counts / names of
test-constraints, test-forsplit, test-aliases).  The reference itself (R) cannot run in this image.
"""
from __future__ import annotations

import copy
import itertools
import warnings

import numpy as np

LD = np.longdouble


class RError(Exception):
    """stop() in the reference."""


def _is_var(x):
    return isinstance(x, lp_var)


def _chr(v):
    """as.character() for set elements."""
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
    """R sum() of doubles: one long-double accumulator walked front to back (src/main/summary.c rsum).
    Written as the plain loop it is (round 2: no code shared with easylp_b200/model.py, which uses a vector cumsum)."""
    acc = LD(0.0)
    for t in np.asarray(v, dtype=float).ravel().tolist():
        acc = acc + LD(t)
    return float(acc)


def _recycle(a, b):
    """R's recycling of two operands, for the two cases the DSL meets: equal lengths, or one scalar."""
    la = [float(t) for t in np.asarray(a, dtype=float).ravel().tolist()]
    lb = [float(t) for t in np.asarray(b, dtype=float).ravel().tolist()]
    if len(la) != len(lb):
        if len(la) == 1:
            la = la * len(lb)
        elif len(lb) == 1:
            lb = lb * len(la)
        else:
            raise RError("longer object length is not a multiple of shorter object length")
    return np.array(la, dtype=float), np.array(lb, dtype=float)


# ------------------------------------------------------------------------------------------------
# Subscripts.  Restated from /root/reference/R/utils.R:108-145 (is_index_valid, find_incorrect_index) and R's own `[`,
# element by element: an index is a list of 1-based numbers or of names; everything is resolved to 0-based positions with
# explicit loops and column-major arithmetic.  (Round 1 shared this block with easylp_b200/model.py; it is now written
# independently so that the differential tests compare two implementations.)
def _index_items(ind):
    """The elements of one subscript as a Python list, and their kind: 'missing', 'num', 'chr' or 'bad'."""
    if ind is None or (isinstance(ind, slice) and ind.start is None and ind.stop is None and ind.step is None):
        return None, "missing"
    if isinstance(ind, (str, np.str_)):
        return [str(ind)], "chr"
    if isinstance(ind, (bool, np.bool_)):
        return None, "bad"
    if isinstance(ind, (int, float, np.integer, np.floating)):
        return [ind], "num"
    if isinstance(ind, slice):
        return None, "bad"
    try:
        items = np.asarray(ind).ravel().tolist() if isinstance(ind, np.ndarray) else list(ind)
    except TypeError:
        return None, "bad"
    flat = []
    for it in items:                                  # one level of nesting is what callers pass (lists, ranges, arrays)
        if isinstance(it, (list, tuple, np.ndarray)):
            flat.extend(np.asarray(it).ravel().tolist())
        else:
            flat.append(it)
    if len(flat) == 0:
        return [], "num"                              # numeric(0) is a valid (empty) numeric subscript: all() of nothing
    if all(isinstance(it, (str, np.str_)) for it in flat):
        return [str(it) for it in flat], "chr"
    if any(isinstance(it, (bool, np.bool_)) for it in flat):
        return None, "bad"
    if all(isinstance(it, (int, float, np.integer, np.floating)) for it in flat):
        return flat, "num"
    return None, "bad"


def _resolve(ind, length, names):
    """is_index_valid + the positions R's `[` would pick; None when the subscript is invalid."""
    items, kind = _index_items(ind)
    if kind == "missing":
        return list(range(length))
    if kind == "num":
        out = []
        for v in items:
            if not (v >= 1 and v < length + 1):       # all(ind >= 1) && all(ind < len + 1)
                return None
            out.append(int(v) - 1)                    # R truncates a fractional subscript towards zero
        return out
    if kind == "chr":
        if names is None:
            return None
        out = []
        for name in items:
            hit = -1
            for pos, cand in enumerate(names):        # match(): the first name that is equal
                if cand == name:
                    hit = pos
                    break
            if hit < 0:
                return None
            out.append(hit)
        return out
    return None


def _positions(shape, dimnames, titles, key):
    """0-based positions per subscript (numpy int arrays), or the reference's error."""
    ndim = len(shape)
    if len(key) == 1:
        total = 1
        for d in shape:
            total *= int(d)
        names = dimnames[0] if (dimnames is not None and ndim == 1) else None
        got = _resolve(key[0], total, names)
        if got is None:
            raise RError("Invalid subscript")
        return [np.array(got, dtype=np.int64)]
    if len(key) != ndim:
        raise RError("Invalid subscript: incorrect number of dimensions")
    out = []
    for d in range(ndim):
        got = _resolve(key[d], int(shape[d]), dimnames[d] if dimnames is not None else None)
        if got is None:
            title = titles[d] if (titles and titles[d]) else d + 1
            raise RError(f"Invalid subscript on dimension '{title}'")
        out.append(np.array(got, dtype=np.int64))
    return out


class Param:
    """parameter(): named array (R/utils.R:356-375), stored as the column-major vector R keeps plus its dim."""

    def __init__(self, a, dimnames):
        self.a = np.asarray(a, dtype=float)
        self.dimnames = dimnames

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def __len__(self):
        return self.a.size

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        shape = self.a.shape
        flat = self.a.flatten(order="F").tolist()         # R's storage order
        try:
            pos = _positions(shape, self.dimnames, None, key)
        except RError:
            raise RError("subscript out of bounds")
        if len(key) == 1:
            picked = [flat[p] for p in pos[0].tolist()]
            dims = [len(picked)]
        else:
            # R's `[` with one index vector per dimension: the result is the outer grid, first subscript fastest
            dims = [len(p) for p in pos]
            strides, acc = [], 1
            for d in shape:
                strides.append(acc)
                acc *= int(d)
            picked = []
            count = 1
            for d in dims:
                count *= d
            for lin in range(count):
                rest, offset = lin, 0
                for d in range(len(dims)):
                    k = rest % dims[d]
                    rest //= dims[d]
                    offset += int(pos[d][k]) * strides[d]
                picked.append(flat[offset])
        if len(picked) == 1:
            return float(picked[0])
        kept = [d for d in dims if d != 1]                # drop = TRUE
        arr = np.array(picked, dtype=float)
        return arr.reshape(kept, order="F") if len(kept) > 1 else arr


def parameter(x, *sets, byrow=False, **named):
    sets = [list(s) for s in sets] + [list(s) for s in named.values()]
    if len(sets) == 0:
        raise RError("Parameter does not have any sets.")
    dims = [len(s) for s in sets]
    total = 1
    for d in dims:
        total *= d
    vals = [float(t) for t in np.asarray(x, dtype=float).ravel().tolist()]
    if len(vals) == 1:
        vals = vals * total
    elif len(vals) != total:
        raise RError("Dimensions of the parameter don't match dimensions of the sets.")
    dn = [[_chr(v) for v in s] for s in sets]
    if byrow:
        if len(sets) != 2:
            raise RError("Use 'byrow = TRUE' only with 2-dimensional arrays.")
        # matrix(x, nrow, byrow = TRUE): x fills row after row
        out = np.empty((dims[0], dims[1]), dtype=float)
        for k, v in enumerate(vals):
            out[k // dims[1], k % dims[1]] = v
        return Param(out, dn)
    # array(x, dim): x fills in storage order, first subscript fastest
    out = np.empty(dims, dtype=float)
    for lin, v in enumerate(vals):
        rest, idx = lin, []
        for d in dims:
            idx.append(rest % d)
            rest //= d
        out[tuple(idx)] = v
    return Param(out, dn)


# ------------------------------------------------------------------------------------------------
class lp_con:
    def __init__(self, mat, dir, rhs):
        self.mat, self.dir, self.rhs = mat, list(dir), np.asarray(rhs, dtype=float)
        self.names, self.rownames = [], []


class ForSplit(list):
    def __init__(self, items, variable, sequence):
        super().__init__(items)
        self.variable, self.sequence = variable, list(sequence)


def for_(body, **index):
    """for(v in seq) body  (R/utils.R:33-64); several indices nest, first outermost."""
    (var, seq), rest = next(iter(index.items())), dict(list(index.items())[1:])
    seq = list(seq)
    if rest:
        return ForSplit([for_(lambda **kw: body(**{var: v}, **kw), **rest) for v in seq], var, seq)
    return ForSplit([body(**{var: v}) for v in seq], var, seq)


class lp_var:
    __array_ufunc__ = None

    def __init__(self, **kw):
        self.__dict__.update(kw)

    # --- R/methods.R:42-47
    def __len__(self):
        return self.coef.shape[0]

    @property
    def dim(self):
        return self.ind.shape if self.has_dim else None

    def copy(self):
        return copy.copy(self)

    # --- `[.lp_var`  R/methods.R:48-69
    def __getitem__(self, key):
        if not self.indexable:
            raise RError("Cannot index this result.")
        if not isinstance(key, tuple):
            key = (key,)
        pos = _positions(self.ind.shape, self.dimnames, self.dimtitles, key)
        x = self.copy()
        old = self.ind.flatten(order="F")
        if len(key) == 1 and self.ind.ndim > 1:
            x.ind = old[pos[0]]
            x.dimnames, x.dimtitles, x.has_dim = None, None, False
        else:
            x.ind = self.ind[np.ix_(*pos)]
            x.dimnames = ([[self.dimnames[d][i] for i in p] for d, p in enumerate(pos)]
                          if self.dimnames is not None else None)
        rows = np.isin(old, x.ind)
        x.raw = False
        x.coef = self.coef[rows, :]
        x.add = self.add[rows]
        return x

    # --- Ops  R/methods.R:114-199
    def _checked(self):
        if np.isnan(self.coef).any() or np.isnan(self.add).any():
            raise RError("Operation resulted in NA values")
        self.raw = False
        return self

    def __pos__(self):
        return self.copy()._checked()

    def __neg__(self):
        x = self.copy()
        x.coef = -x.coef            # `add` is NOT negated (R/methods.R:155-159)
        return x._checked()

    def __mul__(self, k):
        if _is_var(k):
            raise RError("Can't multiply or divide variables in a linear problem")
        x = self.copy()
        kv = _as_vec(k)
        x.coef = horizontal_multiply(x.coef, kv)
        x.add = np.multiply(*_recycle(x.add, kv))
        return x._checked()

    __rmul__ = __mul__

    def __truediv__(self, k):
        if _is_var(k):
            raise RError("Can't multiply or divide variables in a linear problem")
        x = self.copy()
        kv = _as_vec(k)
        x.coef = horizontal_multiply(x.coef, 1.0 / kv)
        a, b = _recycle(x.add, kv)
        x.add = a / b
        return x._checked()

    def __rtruediv__(self, k):
        raise RError("Can't divide by a variable in a linear problem")

    def __add__(self, k):
        x = self.copy()
        if _is_var(k):
            x.coef = horizontal_mat_sum(x.coef, k.coef)
            x.add = np.add(*_recycle(x.add, k.add))
        else:
            x.add = np.add(*_recycle(x.add, _as_vec(k)))
        return x._checked()

    __radd__ = __add__

    def __sub__(self, k):
        x = self.copy()
        if _is_var(k):
            x.coef = horizontal_mat_sum(x.coef, -k.coef)
            a, b = _recycle(x.add, k.add)
            x.add = a - b
        else:
            a, b = _recycle(x.add, _as_vec(k))
            x.add = a - b
        return x._checked()

    def __rsub__(self, k):           # k - x : coef negated, then e2 + e1 (R/methods.R:192-194)
        x = self.copy()
        x.coef = -x.coef
        return x + k

    def __pow__(self, k):
        raise RError("Can't use operations '^', '%%', '%/%' in a linear problem")

    __mod__ = __floordiv__ = __rpow__ = __pow__

    def __invert__(self):            # `!x` (R/methods.R:140-147)
        if not self.binary:
            raise RError("Logical negation '!' only supported on binary variables.")
        return -self + 1

    def __abs__(self):
        raise RError("Function 'abs' is not linear.")

    # --- Compare  R/methods.R:200-225
    def __le__(self, o): return compare(self, "<=", o)
    def __ge__(self, o): return compare(self, ">=", o)
    def __lt__(self, o): return compare(self, "<", o)
    def __gt__(self, o): return compare(self, ">", o)
    def __eq__(self, o): return compare(self, "==", o)
    def __ne__(self, o): raise RError("Inequality '!=' not allowed in linear problems")
    __hash__ = None


def horizontal_multiply(x, mult):     # R/methods.R:82-97
    mult = np.asarray(mult, dtype=float).ravel()
    if x.shape[0] == 1:
        x = np.repeat(x, mult.size, axis=0)
    if mult.size == 1:
        mult = np.repeat(mult, x.shape[0])
    if x.shape[0] != mult.size:
        raise RError("Linear variable must have the same length as multiplier Only values of size one are recycled.")
    with np.errstate(invalid="ignore", over="ignore"):
        return x * mult[:, None]


def horizontal_mat_sum(x, y):         # R/methods.R:98-111
    if x.shape[0] == 1:
        x = np.repeat(x, y.shape[0], axis=0)
    if y.shape[0] == 1:
        y = np.repeat(y, x.shape[0], axis=0)
    if x.shape[0] != y.shape[0]:
        raise RError("Linear variables must have the same length. Only values of size one are recycled.")
    if x.shape != y.shape:
        raise RError("identical(dim(x), dim(y)) is not TRUE")
    return x + y


def compare(e1, op, e2):
    """Compare_lp_var (R/methods.R:200-225) with the reference's operand order."""
    if op == "!=":
        raise RError("Inequality '!=' not allowed in linear problems")
    if not _is_var(e1) and not _is_var(e2):
        raise RError("not a linear comparison")
    if _is_var(e2):
        x = e1 - e2
        rhs = np.zeros(1)
    else:
        x = e1
        rhs = _as_vec(e2)
    if rhs.size == 1:
        rhs = np.repeat(rhs, x.coef.shape[0])
    if rhs.size != x.coef.shape[0]:
        raise RError("length(rhs) == nrow(x$coef) is not TRUE")
    rhs = rhs - x.add
    return lp_con(x.coef, [op] * rhs.size, rhs)


def _colsums(coef):
    """R colSums: per column a sequential long double sum over the rows, rounded once at the end."""
    if coef.shape[0] == 0:
        return np.zeros((1, coef.shape[1]))
    return np.cumsum(coef.astype(LD), axis=0)[-1].astype(float).reshape(1, -1)


def Sum(x, *dots):
    """sum.lp_var (R/methods.R:244-257); falls through to R's sum for plain numbers."""
    if not _is_var(x):
        if any(_is_var(d) for d in dots):
            raise RError("invalid 'type' (list) of argument")
        return _rsum(np.concatenate([_as_vec(x)] + [_as_vec(d) for d in dots]))
    if dots:
        acc = Sum(x)
        for d in dots:
            acc = acc + Sum(d)
        return acc
    x = x.copy()
    x.coef = _colsums(x.coef)
    x.add = np.array([_rsum(x.add)])
    x.indexable = False
    x.raw = False
    return x


def mean(x):
    if not _is_var(x):
        return float(np.mean(_as_vec(x)))
    return Sum(x) / len(x)


def weighted_mean(x, w):
    if len(_as_vec(w)) != len(x):
        raise RError("'x' and 'w' must have the same length")
    return Sum(x * w) / Sum(w)


def cumsum(x):
    """Math.lp_var (R/methods.R:228-242)."""
    if not _is_var(x):
        return np.cumsum(_as_vec(x).astype(LD)).astype(float)
    x = x.copy()
    x.add = np.cumsum(x.add.astype(LD)).astype(float)
    coef = x.coef.copy()
    for i in range(1, coef.shape[0]):
        coef[i, :] = coef[i, :] + coef[i - 1, :]
    x.coef = coef
    x.raw = False
    return x


def sum_for(body, **index):
    """sum_for (R/utils.R:391-411): expand.grid (first index fastest), evaluate, do.call(sum, result)."""
    if not index:
        raise RError("No named indexing variables.")
    names = list(index)
    seqs = [list(index[k]) for k in names]
    result = []
    for cell in itertools.product(*reversed(seqs)):      # last varies slowest
        kw = dict(zip(reversed(names), cell))
        result.append(body(**kw))
    return Sum(*result)


# --- shadowed base functions (R/utils.R:236-333)
def _ensure_not_con(x, fun):
    if isinstance(x, lp_con):
        raise RError(f"Cannot apply function '{fun}' to a constraint.\nDid you accidentally write the constraint inside '{fun}()'?")


def diag(x):
    _ensure_not_con(x, "diag")
    if not _is_var(x):
        return np.diag(np.asarray(x))
    y = x[np.diag(x.ind)]          # ids used as linear subscripts: only valid for the first variable
    y.raw = False
    return y


def apply(X, MARGIN, FUN):
    """base::apply over a linear variable: FUN on every slice that fixes the MARGIN dimensions, results laid out as
    array(dim = dim(X)[MARGIN]) — the first margin runs fastest.  (Independent of easylp_b200/model.py since round 2:
    the cells are walked with an explicit mixed-radix counter.)"""
    _ensure_not_con(X, "apply")
    if not _is_var(X):
        raise RError("apply(): only linear variables are supported by this restatement")
    margins = MARGIN if isinstance(MARGIN, (list, tuple, np.ndarray, range)) else [MARGIN]
    margins = list(margins)
    if any(isinstance(v, (str, np.str_)) for v in margins):
        raise RError("Not all elements of 'MARGIN' are names of dimensions.")
    margins = [int(v) for v in margins]
    shape = list(X.ind.shape)
    for v in margins:
        if v < 1 or v > len(shape):
            raise RError("'MARGIN' does not match dim(X).")
    mdims = [shape[v - 1] for v in margins]
    ncell = 1
    for d in mdims:
        ncell *= d
    coef_rows, adds = [], []
    counter = [1] * len(margins)                       # 1-based position along every margin, first one fastest
    for _ in range(ncell):
        key = [np.arange(1, d + 1) for d in shape]     # a missing subscript = the whole extent
        for v, cpos in zip(margins, counter):
            key[v - 1] = cpos
        z = FUN(X[tuple(key)])
        if z.coef.shape[0] != 1:
            raise RError("number of items to replace is not a multiple of replacement length")
        coef_rows.append(np.array(z.coef[0], dtype=float))
        adds.append(float(z.add[0]))
        for d in range(len(counter)):                  # advance the counter
            counter[d] += 1
            if counter[d] <= mdims[d]:
                break
            counter[d] = 1
    out = X.copy()
    out.ind = np.arange(1, ncell + 1).reshape(mdims, order="F")
    out.dimnames = [X.dimnames[v - 1] for v in margins] if X.dimnames is not None else None
    out.dimtitles = [X.dimtitles[v - 1] for v in margins] if X.dimtitles is not None else None
    out.has_dim = True
    out.coef = np.vstack(coef_rows) if coef_rows else np.zeros((0, X.coef.shape[1]))
    out.add = np.array(adds, dtype=float)
    out.raw = False
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
def name_constraint(con, name):       # R/utils.R:154-165
    n = con.mat.shape[0]
    if not name:
        con.names = [""] * n
        con.rownames = [""] * n
        return con
    con.names = [name] * n
    con.rownames = [f"{name}[{k}]" for k in range(1, n + 1)] if n > 1 else [name]
    return con


def flatten_for_split(split, init_name=""):    # R/utils.R:66-94
    atoms = []

    def add(x, name):
        if isinstance(x, ForSplit):
            name = name.replace("]", ",", 1)
            for k, item in enumerate(x):
                add(item, f"{name}{x.variable}={_chr(x.sequence[k])}]")
        else:
            if isinstance(x, lp_con):
                x = name_constraint(x, name)
                x.names = [init_name] * x.mat.shape[0]
            atoms.append((name, x))

    add(split, (init_name or "") + "[")
    return atoms


class easylp:
    """R6 class `easylp` (R/class.R:51-648), dense like the reference."""

    def __init__(self):
        self.variables = {}
        self.aliases = {}
        self.mat = np.zeros((0, 0))
        self.dir, self.rhs, self.names, self.rownames = [], np.zeros(0), [], []
        self.objective_fun = np.zeros(0)
        self.objective_add = 0.0
        self.n_var = 0
        self.direction = "min"

    def var(self, name, *sets, integer=False, binary=False, lower=-np.inf, upper=np.inf, **named):   # R/class.R:85-179
        if name in self.variables:
            raise RError(f"Variable '{name}' already defined in this model.")
        if lower > upper:
            warnings.warn("Lower bound is higher than upper bound. Problem will be unfeasible.")
        if binary:
            integer = False
            lower, upper = 0.0, 1.0
        titles = [""] * len(sets) + list(named)
        sets = [list(s) for s in sets] + [list(s) for s in named.values()]
        if not sets:
            sets, titles = [[""]], ["scalar"]
        dims = tuple(len(s) for s in sets)
        ln = int(np.prod(dims))
        ind = (np.arange(1, ln + 1) + self.n_var).reshape(dims, order="F")
        coef = np.hstack([np.zeros((ln, self.n_var)), np.eye(ln)])
        for v in self.variables.values():
            v.coef = np.hstack([v.coef, np.zeros((v.coef.shape[0], ln))])
        for key, v in list(self.aliases.items()):
            if _is_var(v):
                v.coef = np.hstack([v.coef, np.zeros((v.coef.shape[0], ln))])
        self.mat = np.hstack([self.mat, np.zeros((self.mat.shape[0], ln))])
        self.objective_fun = np.concatenate([self.objective_fun, np.zeros(ln)])
        x = lp_var(name=name, ind=ind, dimnames=[[_chr(v) for v in s] for s in sets], dimtitles=titles, has_dim=True,
                   type="integer" if integer else ("binary" if binary else "real"), integer=integer, binary=binary,
                   bound=[float(lower), float(upper)], indexable=True, raw=True, coef=coef, add=np.zeros(ln))
        self.variables[name] = x
        self.n_var += ln
        return x

    def alias(self, **named):
        self.aliases.update(named)

    def _join(self, con):            # join_constraints (R/utils.R:95-106)
        self.mat = np.vstack([self.mat, con.mat])
        self.dir += list(con.dir)
        self.rhs = np.concatenate([self.rhs, con.rhs])
        self.names += list(con.names)
        self.rownames += list(con.rownames)

    def con(self, *unnamed, **named):   # R/class.R:189-220
        items = [(None, c) for c in unnamed] + list(named.items())
        for k, (name, c) in enumerate(items, 1):
            ref = name or k
            if callable(c):
                try:
                    c = c()
                except Exception as e:
                    raise RError(f"Constraint '{ref}' evaluated to an error:\n{e}") from e
            if isinstance(c, ForSplit):
                split = flatten_for_split(c, name or "")
                if not split or not isinstance(split[0][1], lp_con):
                    raise RError("Constraint did not evaluate to an (in)equality.")
                for _, atom in split:
                    self._join(atom)
                continue
            if not isinstance(c, lp_con):
                raise RError(f"Constraint '{ref}' did not evaluate to an (in)equality.")
            if c.mat.shape[0] == 0:
                warnings.warn(f"Constraint '{ref}' is empty.")
                continue
            self._join(name_constraint(c, name))
        return self

    def associate(self, x, binary, max1=None, max0=None, min1=None, min0=None):   # R/class.R:332-358
        """Dense restatement for a plain variable `x` (update_bounds of a variable is its own bound pair)."""
        lo, up = x.bound
        max1 = up if max1 is None else max1
        max0 = lo if max0 is None else max0
        min1 = lo if min1 is None else min1
        min0 = lo if min0 is None else min0
        if not all(np.isfinite(v) for v in (max1, max0, min1, min0)):
            raise RError("is.finite(max1), is.finite(max0), is.finite(min1), is.finite(min0) are not all TRUE")
        if not binary.binary:
            warnings.warn("Variable is not binary.")
        if max1 != up or max0 != up:
            self.con(assoc_max=x <= max0 + (max1 - max0) * binary)
        if min1 != lo or min0 != lo:
            self.con(assoc_min=x >= min0 + (min1 - min0) * binary)
        return self

    def uncon(self, name):           # R/class.R:308-316
        names = [name] if isinstance(name, str) else list(name)
        keep = np.array([n not in names for n in self.names], dtype=bool)
        self.mat = self.mat[keep]
        self.dir = [d for d, k in zip(self.dir, keep) if k]
        self.rhs = self.rhs[keep]
        # the reference does not shrink `names` (R/class.R:311-314 touch mat, dir, rhs only)
        self.rownames = [r for r, k in zip(self.rownames, keep) if k]
        return self

    def _objective(self, x, direction):   # R/class.R:509-531
        self.direction = direction
        if isinstance(x, lp_con):
            raise RError("Objective function evaluated to a constraint. It must evaluate to a variable or sum of variables.")
        if not _is_var(x):
            raise RError("Objective function didn't evaluate to a variable or sum of variables.")
        if len(x) == 0:
            raise RError("Objective function doesn't contain any variables.")
        if len(x) > 1:
            raise RError("Objective function contains multiple variables. Please wrap them in a sum().")
        self.objective_fun = x.coef[0].copy()
        self.objective_add = float(x.add[0])
        return self.objective_fun

    def min(self, objective, transform=None):
        return self._objective(objective, "min")

    def max(self, objective, transform=None):
        return self._objective(objective, "max")

    # ---- what `$solve()` hands to lp_solve (R/class.R:260-274), as canonical CSR + vectors -------
    def canonical(self):
        m, n = self.mat.shape[0], self.n_var
        mat = self.mat
        row_ptr = np.zeros(m + 1, np.int32)
        cols, vals = [], []
        for i in range(m):
            nz = np.nonzero(mat[i] != 0)[0]
            cols.append(nz.astype(np.int32))
            vals.append(mat[i, nz])
            row_ptr[i + 1] = row_ptr[i] + nz.size
        lb = np.concatenate([np.repeat(v.bound[0], v.ind.size) for v in self.variables.values()]) if n else np.zeros(0)
        ub = np.concatenate([np.repeat(v.bound[1], v.ind.size) for v in self.variables.values()]) if n else np.zeros(0)
        sense = np.array([{"<=": 0, "<": 0, ">=": 1, ">": 1, "==": 2}[d] for d in self.dir], dtype=np.int8)
        return dict(m=m, n=n, row_ptr=row_ptr,
                    col_idx=np.concatenate(cols) if cols else np.zeros(0, np.int32),
                    vals=np.concatenate(vals) if vals else np.zeros(0),
                    dir=list(self.dir), sense=sense, rhs=self.rhs.copy(), c=self.objective_fun.copy(),
                    objective_add=self.objective_add, lb=lb, ub=ub, maximize=self.direction == "max",
                    names=list(self.names), rownames=list(self.rownames),
                    is_integer=np.concatenate([np.repeat(bool(v.integer or v.binary), v.ind.size)
                                               for v in self.variables.values()]).astype(np.uint8) if n else np.zeros(0, np.uint8))
