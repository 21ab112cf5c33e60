"""Symbolic lowering of `for` / `sum_for` bodies to index-set descriptors.

The reference evaluates the body of `for (v in seq) body` once per atom (R/utils.R:50-53) and the body of `sum_for`
once per grid row (R/utils.R:405-408); each evaluation indexes variables (`[.lp_var`, R/methods.R:48-69) and runs the
arithmetic of R/methods.R:148-199.  That interpreter loop dominates model-build time once the solver is fast.  Here
the body is evaluated ONCE with symbolic loop indices.  When it is affine in fully indexed raw variables the trace
yields, per body term, a *family*: a loop nest, a column-offset table per loop and a coefficient table over the loops
the coefficient depends on.  `easylp._csr()` ships the families to `elp_assemble_lowered`, which expands them into the
term stream on the device and folds it (include/easylp_abi.h).  R would do the same with `substitute()`.

Exactness.  The lowered stream must fold to the same doubles as the reference's evaluation order, so the trace keeps
the reference's fold structure explicit:
  * a GROUP is one `sum_for` (or one bare expression): its terms are added left to right in grid order (first name
    fastest, R/utils.R:402), exactly the pending `Reduce('+')` of R/methods.R:248-250;
  * a scalar applied to a `sum_for` result multiplies the FOLDED sum (R/methods.R:82-97 runs on the evaluated
    argument): it is recorded as a post-fold multiplier of the group, not pushed into the terms;
  * `e1 + e2`, `e1 - e2` add folded groups left to right (R/methods.R:98-111);
  * constants follow R/methods.R:155-164,192-194 (`-x` leaves `add` alone, `k - x` adds k, `x / k` is `coef * (1/k)`).
Anything the trace cannot prove equal to the eager evaluation raises NotLowerable (or any other exception) inside the
trace and the caller silently takes the eager path — which also reports the user's errors at the usual place.
Coefficient expressions are evaluated with numpy element-wise float64 operations, which are the same IEEE operations
the per-cell Python floats of the eager path perform.
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

_I = np.int32
MAX_LOOPS = 6
MAX_GROUP_MUL = 4
_uid = itertools.count(1)
_depth = 0          # > 0 while the body of an enclosing for / sum_for is being traced


class NotLowerable(Exception):
    pass


def _refuse(*_a, **_k):
    raise NotLowerable()


def _is_number(k):
    return isinstance(k, (int, float, np.integer, np.floating)) and not isinstance(k, (bool, np.bool_))


# ---- symbolic scalars ------------------------------------------------------------------------------
class Coef:
    """A numeric scalar that depends on loop variables: constants, loop values, parameter look-ups, + - * /."""
    __array_ufunc__ = None
    __hash__ = None

    def __init__(self, kind, a=None, b=None, op=None):
        self.kind, self.a, self.b, self.op = kind, a, b, op

    @staticmethod
    def wrap(k):
        if isinstance(k, Coef):
            return k
        if _is_number(k):
            return Coef("const", float(k))
        raise NotLowerable()

    def _bin(self, op, other, swap=False):
        if isinstance(other, (SymExpr, SymCon)):
            return NotImplemented
        o = Coef.wrap(other)
        return Coef("bin", o, self, op) if swap else Coef("bin", self, o, op)

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __neg__(self): return Coef("neg", self)
    def __pos__(self): return self

    # comparisons against symbolic variables are reflected to them (as float.__le__(lp_var) is in the eager path)
    def _cmp(self, o):
        if isinstance(o, SymExpr):
            return NotImplemented
        raise NotLowerable()
    __le__ = __ge__ = __lt__ = __gt__ = __eq__ = __ne__ = _cmp
    # anything that needs the VALUE of the scalar cannot be traced
    __bool__ = __int__ = __float__ = __index__ = __len__ = __iter__ = __str__ = __format__ = _refuse
    __pow__ = __rpow__ = __mod__ = __rmod__ = __floordiv__ = __rfloordiv__ = __abs__ = _refuse

    def syms(self, out):
        if self.kind == "idx":
            out[self.a.uid] = self.a
        elif self.kind == "param":
            for s in self.b:
                if not isinstance(s, (int, np.integer)):
                    out[s[0].uid] = s[0]
        elif self.kind == "sumover":
            inner = {}
            self.a.syms(inner)
            for u, s in inner.items():
                if all(u != l.uid for l in self.b):
                    out[u] = s
        else:
            for c in (self.a, self.b):
                if isinstance(c, Coef):
                    c.syms(out)
        return out

    def as_subscript(self):
        """(loop variable, integer offset) when the scalar is `v`, `v + k` or `v - k`"""
        if self.kind == "idx":
            return self.a, 0
        if self.kind == "bin" and self.op in "+-":
            a, b = self.a, self.b
            if a.kind == "idx" and b.kind == "const" and float(b.a).is_integer():
                return a.a, int(b.a) if self.op == "+" else -int(b.a)
            if self.op == "+" and b.kind == "idx" and a.kind == "const" and float(a.a).is_integer():
                return b.a, int(a.a)
        raise NotLowerable()


class SymIndex(Coef):
    """The loop variable of a `for` / `sum_for` during a trace."""

    def __init__(self, name, seq):
        super().__init__("idx", self)
        self.name, self.seq, self.uid = name, list(seq), next(_uid)
        if not self.seq:
            raise NotLowerable()
        if len(self.seq) > 256 and type(self.seq[0]) in (int, float) and np.asarray(self.seq).dtype.kind in "iuf":
            self.numeric = True                         # long numeric ranges: one vectorised check instead of a Python loop
        else:
            self.numeric = all(_is_number(v) for v in self.seq)
            if not self.numeric and not all(isinstance(v, (str, np.str_)) for v in self.seq):
                raise NotLowerable()


def _evaluate(c: Coef, axes: dict, ndim: int, expand=None) -> np.ndarray:
    """value of the scalar over the loop grid: shape has the loop's extent on the axes it depends on, 1 elsewhere.
    `expand[uid]`: positions of that loop's values along its axis when the axis is a fused (parent, ragged) loop."""
    one = (1,) * ndim
    expand = expand or {}
    if c.kind == "const":
        return np.full(one, c.a)
    if c.kind == "idx":
        s = c.a
        if not s.numeric or s.uid not in axes:
            raise NotLowerable()
        vals = np.asarray(s.seq, dtype=float)
        if s.uid in expand:
            vals = vals[expand[s.uid]]
        shape = list(one)
        shape[axes[s.uid]] = vals.size
        return vals.reshape(shape)
    if c.kind == "param":
        idx = []
        for s in c.b:
            if isinstance(s, (int, np.integer)):
                idx.append(np.full(one, int(s), dtype=np.int64))
            else:
                sym, tab = s
                if sym.uid not in axes:
                    raise NotLowerable()
                t = tab.astype(np.int64)
                if sym.uid in expand:
                    t = t[expand[sym.uid]]
                shape = list(one)
                shape[axes[sym.uid]] = t.size
                idx.append(t.reshape(shape))
        return np.asarray(c.a[tuple(np.broadcast_arrays(*idx))], dtype=float)
    if c.kind == "neg":
        return -_evaluate(c.a, axes, ndim, expand)
    if c.kind == "bin":
        a, b = _evaluate(c.a, axes, ndim, expand), _evaluate(c.b, axes, ndim, expand)
        with np.errstate(all="ignore"):
            return a + b if c.op == "+" else a - b if c.op == "-" else a * b if c.op == "*" else a / b
    if c.kind == "sumover":
        # `Sum(*cells)`: add <- add + cell.add, cell after cell in double (model.Sum); first name fastest = last axis
        loops = c.b                                     # slowest first: they become the trailing axes
        if expand or any(getattr(l, "ragged", None) for l in loops):
            raise NotLowerable()                        # constants summed over a ragged set: not traced
        ax2 = dict(axes)
        for i, l in enumerate(loops):
            ax2[l.uid] = ndim + i
        v = _evaluate(c.a, ax2, ndim + len(loops))
        full = list(v.shape)
        for i, l in enumerate(loops):
            full[ndim + i] = len(l.seq)
        v = np.broadcast_to(v, full)
        return np.asarray(np.cumsum(v.reshape(v.shape[:ndim] + (-1,)), axis=-1)[..., -1], dtype=float)
    raise NotLowerable()


def _subscript_table(length, names, sym: SymIndex, off: int):
    """0-based positions of the loop's values in one dimension — `[`'s own rules (model._positions)"""
    from . import model
    if off and not sym.numeric:
        raise NotLowerable()
    vals = [v + off for v in sym.seq] if off else sym.seq
    try:
        return model._positions((length,), [names] if names is not None else None, None, (vals,))[0].astype(_I)
    except model.EasyLpError:
        raise NotLowerable()


def param_getitem(p, key):
    """`cost[f, m]` with symbolic subscripts"""
    from . import model
    a = p.a
    if len(key) != a.ndim:
        raise NotLowerable()
    subs = []
    for d, k in enumerate(key):
        names = p.dimnames[d] if p.dimnames is not None else None
        if isinstance(k, Coef):
            sym, off = k.as_subscript()
            subs.append((sym, _subscript_table(a.shape[d], names, sym, off)))
        elif _is_number(k) or isinstance(k, (str, np.str_)):
            try:
                subs.append(int(model._positions((a.shape[d],), [names] if names is not None else None, None, (k,))[0][0]))
            except model.EasyLpError:
                raise NotLowerable()
        else:
            raise NotLowerable()
    return Coef("param", a, subs)


# ---- symbolic one-row affine expressions -------------------------------------------------------------
class Term:
    __slots__ = ("block", "col0", "tabs", "coef")

    def __init__(self, block, col0, tabs, coef):
        self.block, self.col0, self.tabs, self.coef = block, col0, tabs, coef     # tabs: uid -> (sym, int32 offsets)


class Group:
    __slots__ = ("loops", "terms", "post", "from_slice")

    def __init__(self, loops, terms, post):
        self.loops, self.terms, self.post = tuple(loops), list(terms), list(post)
        self.from_slice = False      # a canonical row — `sum(x[f, ])`, a row of an alias: every column occurs at most once

    def mapped(self, f):
        g = Group(self.loops, [Term(t.block, t.col0, t.tabs, f(t.coef)) for t in self.terms], self.post)
        g.from_slice = self.from_slice
        return g


class SymExpr:
    """One row `coef . x + add` whose entries depend on loop variables (the symbolic twin of model.lp_var)."""
    __array_ufunc__ = None
    __hash__ = None

    def __init__(self, groups, add=None):
        self.groups, self.add = list(groups), add

    def _add(self):
        return self.add if self.add is not None else Coef("const", 0.0)

    # Arith_lp_var, R/methods.R:148-199
    def __pos__(self):
        return self

    def __neg__(self):                                  # coef negated, `add` kept (R/methods.R:155-159)
        return SymExpr([g.mapped(lambda c: Coef("neg", c)) for g in self.groups], self.add)

    def _scale(self, k, add_op):
        if isinstance(k, (SymExpr, SymCon)):
            raise NotLowerable()                        # var * var: the eager path raises the reference's error
        k = Coef.wrap(k)
        if len(self.groups) != 1:
            raise NotLowerable()                        # (a + b) * k scales the folded sum of both: three levels
        g = self.groups[0]
        if ((not g.loops and len(g.terms) == 1) or g.from_slice) and not g.post:
            ng = g.mapped(lambda c: Coef("bin", c, k, "*"))                     # coef * k, entry by entry (no entry repeats)
        else:
            ng = Group(g.loops, g.terms, g.post + [k])                          # multiplies the folded sum
            if len(ng.post) > MAX_GROUP_MUL:
                raise NotLowerable()
        return SymExpr([ng], add_op(self._add(), k) if (self.add is not None) else None)

    def __mul__(self, k):
        return self._scale(k, lambda a, kk: Coef("bin", a, kk, "*"))
    __rmul__ = __mul__

    def __truediv__(self, k):                           # coef * (1/k), add / k  (R/methods.R:162-164)
        if isinstance(k, (SymExpr, SymCon)):
            raise NotLowerable()
        k = Coef.wrap(k)
        inv = Coef("bin", Coef("const", 1.0), k, "/")
        return self._scale(inv, lambda a, _kk: Coef("bin", a, k, "/"))

    __rtruediv__ = __pow__ = __rpow__ = __mod__ = __floordiv__ = __invert__ = __abs__ = _refuse
    __bool__ = __iter__ = __getitem__ = _refuse

    def __len__(self):
        return 1

    def _plus(self, o, sign):
        if isinstance(o, SymExpr):
            # a + (b + c) adds the FOLDED pair (b + c).  That is the left fold a + b + c only if no entry receives both
            # b and c, which is certain when they sit on different variables (disjoint column ranges).
            blocks = [{t.block for t in g.terms} for g in o.groups]
            if len(blocks) > 1 and any(None in b_ for b_ in blocks):
                raise NotLowerable()                    # rows of aliases may share columns with anything
            if any(blocks[i] & blocks[j] for i in range(len(blocks)) for j in range(i)):
                raise NotLowerable()
            og = o.groups if sign > 0 else [g.mapped(lambda c: Coef("neg", c)) for g in o.groups]
            add = None
            if self.add is not None or o.add is not None:
                add = Coef("bin", self._add(), o._add(), "+" if sign > 0 else "-")
            return SymExpr(self.groups + list(og), add)
        k = Coef.wrap(o)
        return SymExpr(self.groups, Coef("bin", self._add(), k, "+" if sign > 0 else "-"))

    def __add__(self, o): return self._plus(o, +1)
    def __radd__(self, o): return self._plus(o, +1)
    def __sub__(self, o): return self._plus(o, -1)

    def __rsub__(self, k):                              # k - x: coef negated, then `+ k` (R/methods.R:192-194)
        return (-self)._plus(k, +1)

    # Compare_lp_var, R/methods.R:200-225
    def _cmp(self, op, o):
        if isinstance(o, SymExpr):
            x = self - o
            rhs = Coef("const", 0.0)
        else:
            x, rhs = self, Coef.wrap(o)
        if x.add is not None:
            rhs = Coef("bin", rhs, x.add, "-")
        return SymCon(x.groups, op, rhs)

    def __le__(self, o): return self._cmp("<=", o)
    def __ge__(self, o): return self._cmp(">=", o)
    def __lt__(self, o): return self._cmp("<", o)
    def __gt__(self, o): return self._cmp(">", o)
    def __eq__(self, o): return self._cmp("==", o)
    def __ne__(self, o): raise NotLowerable()

    def syms(self):
        """free loop variables (a group's own `sum_for` loops are bound)"""
        out = {}
        for g in self.groups:
            inner = {}
            for t in g.terms:
                for u, (s, _tab) in t.tabs.items():
                    inner[u] = s
                t.coef.syms(inner)
            bound = {l.uid for l in g.loops}
            out.update({u: s for u, s in inner.items() if u not in bound})
            for k in g.post:
                k.syms(out)
        if self.add is not None:
            self.add.syms(out)
        return out


class SymCon:
    def __init__(self, groups, op, rhs):
        self.groups, self.op, self.rhs = groups, op, rhs
        self.row_loops = []          # vector atoms: the loops over the rows of ONE atom, slowest first
    __bool__ = _refuse


class SymVec:
    """`x[f, ]`: the rows of a raw variable along its sliced dimensions (one entry with coefficient 1 per row).  Only
    what `sum(x[f, ])`, `mean(x[f, ])` and a scalar factor need is traced."""
    __array_ufunc__ = None
    __hash__ = None

    def __init__(self, term: Term, implicit):
        self.term, self.implicit = term, list(implicit)      # implicit loops in dimension order (first fastest)

    def __len__(self):
        return int(np.prod([len(l.seq) for l in self.implicit]))

    def _scaled(self, k):
        if not isinstance(k, Coef) and not _is_number(k):
            raise NotLowerable()                             # vectors recycle row-wise: not traced
        k = Coef.wrap(k)
        t = self.term
        return SymVec(Term(t.block, t.col0, t.tabs, Coef("bin", t.coef, k, "*")), self.implicit)

    def __mul__(self, k): return self._scaled(k)
    __rmul__ = __mul__

    def __truediv__(self, k):
        if not isinstance(k, Coef) and not _is_number(k):
            raise NotLowerable()
        return self._scaled(Coef("bin", Coef("const", 1.0), Coef.wrap(k), "/"))

    def __neg__(self):
        t = self.term
        return SymVec(Term(t.block, t.col0, t.tabs, Coef("neg", t.coef)), self.implicit)

    def as_sum(self):
        """sum.lp_var of the rows = colSums: every column occurs in one row only, so each column sum is its entry"""
        g = Group(tuple(reversed(self.implicit)), [self.term], [])
        g.from_slice = True
        return SymExpr([g])

    __add__ = __radd__ = __sub__ = __rsub__ = __rtruediv__ = __pow__ = __getitem__ = __iter__ = __bool__ = _refuse

    def _cmp(self, op, o):
        """`x[, b, 2] >= 1`: one row per entry of the slice (R/methods.R:200-225 with a recycled scalar rhs)"""
        if not isinstance(o, Coef) and not _is_number(o):
            raise NotLowerable()                             # vector right-hand sides: not traced
        con = SymCon([Group((), [self.term], [])], op, Coef.wrap(o))
        con.row_loops = list(reversed(self.implicit))        # rows of the atom: first sliced dimension fastest
        return con

    def __le__(self, o): return self._cmp("<=", o)
    def __ge__(self, o): return self._cmp(">=", o)
    def __lt__(self, o): return self._cmp("<", o)
    def __gt__(self, o): return self._cmp(">", o)
    def __eq__(self, o): return self._cmp("==", o)
    def __ne__(self, o): raise NotLowerable()


def sym_sum(x, *dots):
    """sum.lp_var (R/methods.R:244-257) on traced arguments: every argument is folded on its own, then `Reduce('+')`"""
    parts = []
    for a in (x,) + tuple(dots):
        if isinstance(a, SymVec):
            parts.append(a.as_sum())
        elif isinstance(a, SymExpr) or isinstance(a, Coef) or _is_number(a):
            parts.append(a)
        else:
            raise NotLowerable()
    if not isinstance(parts[0], SymExpr):
        raise NotLowerable()
    acc = parts[0]
    for p_ in parts[1:]:
        acc = acc._plus(p_, +1)
    return acc


def is_traced(*xs):
    return any(isinstance(x, (SymVec, SymExpr)) for x in xs)


def var_getitem(x, key):
    """`x[s, t]` with symbolic subscripts (R/methods.R:48-69).  On a raw variable block: one entry with coefficient 1,
    or — with empty subscripts, `x[f, ]` — a SymVec.  On any other indexable variable (an alias, `rowSums(x)`, ...): the
    row picked by ONE loop variable, as a fused (loop value, entry of that row) family."""
    from . import model
    if not (x.indexable and x.has_dim and x.ind.size == x.nrow):
        raise NotLowerable()
    ind = x.ind
    if len(key) != ind.ndim:
        raise NotLowerable()
    raw = bool(getattr(x, "raw", False)) and np.array_equal(x.t_col, ind.flatten(order="F") - 1)
    base = int(ind.flat[0]) - 1 if raw else 0
    stride, tabs, implicit = 1, {}, []
    for d, k in enumerate(key):
        names = x.dimnames[d] if x.dimnames is not None else None
        if isinstance(k, Coef):
            sym, off = k.as_subscript()
            tab = _subscript_table(ind.shape[d], names, sym, off).astype(np.int64) * stride
            if sym.uid in tabs:
                tab = tabs[sym.uid][1] + tab
            tabs[sym.uid] = (sym, tab)
        elif k is None or (isinstance(k, slice) and k == slice(None)):
            if not raw:
                raise NotLowerable()
            sym = SymIndex(f"_dim{d + 1}", range(1, ind.shape[d] + 1))
            tabs[sym.uid] = (sym, np.arange(ind.shape[d], dtype=np.int64) * stride)
            implicit.append(sym)
        elif _is_number(k) or isinstance(k, (str, np.str_)):
            try:
                base += stride * int(model._positions((ind.shape[d],), [names] if names is not None else None, None, (k,))[0][0])
            except model.EasyLpError:
                raise NotLowerable()
        else:
            raise NotLowerable()                        # vectors of subscripts: not traced
        stride *= ind.shape[d]
    if len(tabs) == len(implicit):
        raise NotLowerable()                            # no loop variable involved: the eager `[` answers
    if raw:
        tabs = {u: (sy, t.astype(_I)) for u, (sy, t) in tabs.items()}
        term = Term(id(x), base, tabs, Coef("const", 1.0))
        return SymVec(term, implicit) if implicit else SymExpr([Group((), [term], [])])
    # a row of a canonical multi-entry variable, chosen by one loop variable
    if len(tabs) != 1:
        raise NotLowerable()
    (parent, rowtab), = tabs.values()
    rowpos = rowtab + base
    xc = x.copy()._canon()
    ptr = np.asarray(xc._row_ptr(), dtype=np.int64)
    counts = ptr[rowpos + 1] - ptr[rowpos]
    total = int(counts.sum())
    if total == 0:
        raise NotLowerable()
    src = np.repeat(ptr[rowpos] - np.r_[0, np.cumsum(counts)[:-1]], counts) + np.arange(total, dtype=np.int64)
    e = SymIndex("_entry", range(total))
    e.ragged = (parent, counts)
    term = Term(None, 0, {e.uid: (e, np.asarray(xc.t_col)[src].astype(_I))},
                Coef("param", np.asarray(xc.t_val, dtype=float)[src], [(e, np.arange(total, dtype=_I))]))
    add = None
    if np.any(np.asarray(x.add) != 0.0):
        add = Coef("param", np.asarray(x.add, dtype=float), [(parent, rowpos.astype(_I))])
    g = Group((e,), [term], [])
    g.from_slice = True
    return SymExpr([g], add)


def has_symbolic(key):
    return any(isinstance(k, Coef) for k in key)


# ---- ragged index sets ---------------------------------------------------------------------------------
class RaggedSeq:
    """`out[[v]]` with a symbolic v: one list of values per value of the enclosing loop variable"""

    def __init__(self, parent: SymIndex, lists):
        self.parent, self.lists = parent, [list(x) for x in lists]


class IndexSets:
    """A list of index sets keyed by the values of another index (R: a named list, `out[[v]]`), e.g. the arcs that
    leave node v.  `sets[v]` is the plain list when v is a value; inside a for/sum_for trace it stands for all of
    them at once, so `for (v in V) sum_for(a = out[[v]], ...)` lowers to one family over the (v, a) pairs."""

    def __init__(self, mapping):
        self.map = {k: list(v) for k, v in dict(mapping).items()}

    def __getitem__(self, key):
        if isinstance(key, Coef):
            sym, off = key.as_subscript()
            if off:
                raise NotLowerable()
            try:
                return RaggedSeq(sym, [self.map[v] for v in sym.seq])
            except (KeyError, TypeError):
                raise NotLowerable()
        return self.map[key]

    def __len__(self):
        return len(self.map)


# ---- sum_for -----------------------------------------------------------------------------------------
def _sum_group(cell: SymExpr, loops):
    """`do.call(sum, cells)`: every cell is folded on its own first (sum.lp_var), then the cells are added in grid
    order.  With one bare group whose entries sit on distinct variables the per-cell fold is the identity."""
    if len(cell.groups) == 1 and getattr(cell.groups[0], "from_slice", False) and not cell.groups[0].post:
        # the cell is `sum` of a slice (`tdm[, m] * k[m]`): its own fold is the identity (every column once), so the
        # slice's rows simply become the fastest loops of this sum
        g = cell.groups[0]
        add = Coef("sumover", cell.add, tuple(loops)) if cell.add is not None else None
        return SymExpr([Group(tuple(loops) + g.loops, g.terms, [])], add)
    if any(g.loops or g.post for g in cell.groups):
        raise NotLowerable()                            # a sum of sums folds three levels deep
    terms = [t for g in cell.groups for t in g.terms]   # `a + b` inside the cell: entries of different variables
    if len({t.block for t in terms}) != len(terms):
        raise NotLowerable()                            # the same variable twice in a cell: entries may collide
    add = Coef("sumover", cell.add, tuple(loops)) if cell.add is not None else None
    return SymExpr([Group(loops, terms, [])], add)


def try_sum_for(body, index):
    """Returns a SymExpr (inside an enclosing trace), a materialised lp_var (top level), or None = take the eager path."""
    names = list(index)
    try:
        syms = {}
        for k in names:
            seq = index[k]
            if isinstance(seq, RaggedSeq):              # one set per value of an enclosing loop: a fused (parent, k) loop
                if any(len(x) == 0 for x in seq.lists):
                    return None                         # an empty sum is an error in the eager path
                syms[k] = SymIndex(k, [v for x in seq.lists for v in x])
                syms[k].ragged = (seq.parent, np.asarray([len(x) for x in seq.lists], dtype=np.int64))
            else:
                syms[k] = SymIndex(k, seq)
        global _depth
        _depth += 1
        try:
            cell = body(**syms)
        finally:
            _depth -= 1
        if isinstance(cell, SymVec):                    # `sum_for(m = M, tdm[, m] * k[m])`: every cell is summed first
            cell = cell.as_sum()
        if not isinstance(cell, SymExpr):
            return None
        loops = [syms[k] for k in reversed(names)]      # slowest first; the first name is the fastest (R/utils.R:402)
        if len(loops) > MAX_LOOPS:
            return None
        expr = _sum_group(cell, loops)
        if _depth > 0:
            return expr                                 # inside an enclosing trace (even if no outer index is used here)
        free = expr.syms()
        if any(u not in {l.uid for l in loops} for u in free):
            return None                                 # a loop variable that nobody binds: cannot happen at top level
        if any(getattr(l, "ragged", None) for l in expr.groups[0].loops):
            return None                                 # ragged pairs are only expanded on the device
        first = body(**{k: index_first(index[k]) for k in names})           # metadata of the eager result
        return _materialise(expr, list(expr.groups[0].loops), first)
    except Exception:
        return None


def index_first(seq):
    return list(seq)[0]


def _materialise(expr: SymExpr, loops, first_cell):
    """Top-level `sum_for`: the pending term list the eager loop would have produced, emitted in one vectorised pass."""
    from . import model
    out = model.Sum(first_cell)
    axes = {l.uid: i for i, l in enumerate(loops)}
    ext = [len(l.seq) for l in loops]
    g = expr.groups[0]
    cols, vals = [], []
    for t in g.terms:
        c = np.full(ext, t.col0, dtype=np.int64)
        for u, (_s, tab) in t.tabs.items():
            shape = [1] * len(loops)
            shape[axes[u]] = ext[axes[u]]
            c = c + tab.reshape(shape)
        v = np.broadcast_to(_evaluate(t.coef, axes, len(loops)), ext)
        if not np.all(np.isfinite(v)):
            raise NotLowerable()
        cols.append(c.reshape(-1))
        vals.append(v.reshape(-1))
    col = np.stack(cols, axis=1).reshape(-1).astype(_I)            # cell-major, body order inside a cell
    val = np.stack(vals, axis=1).reshape(-1).astype(float)
    keep = val != 0.0                                              # `_scaled` keeps no zero entries
    col, val = col[keep], val[keep]
    add = float(_evaluate(expr._add(), axes, len(loops)).reshape(-1)[0])
    if np.isnan(add):
        raise NotLowerable()
    out.t_row, out.t_col, out.t_val = np.zeros(col.size, _I), col, val
    out.canonical, out._ptr = False, None
    if int(np.prod(ext)) == 1:
        out._canon()
    out.add = np.array([add])
    return out._checked()


# ---- for --------------------------------------------------------------------------------------------
class LoweredFor:
    """`for (v in seq) body` whose body traced to one symbolic (in)equality: loops outermost first."""

    def __init__(self, loops, con: SymCon):
        self.loops, self.con = list(loops), con


def try_for(body, index):
    names = list(index)
    try:
        syms = {k: SymIndex(k, index[k]) for k in names}
        global _depth
        _depth += 1
        try:
            r = body(**syms)
        finally:
            _depth -= 1
        loops = [syms[k] for k in names]                # first index outermost (model.for_)
        if isinstance(r, LoweredFor):
            loops, r = loops + r.loops, r.con
        if not isinstance(r, SymCon) or len(loops) > MAX_LOOPS:
            return None
        low = LoweredFor(loops, r)
        free = SymExpr(r.groups, r.rhs).syms()
        if all(u in {l.uid for l in loops + r.row_loops} for u in free):
            low.block = build_block(low)                # descriptors now, so failures fall back here
        elif _depth == 0:
            return None                                 # a loop variable nobody binds, at top level: not ours
        return low
    except Exception:
        return None


class LoweredCon:
    """A block of constraint rows kept as families (the lowered twin of model.lp_con)."""

    def __init__(self, nrow, dir_, rhs, families, groups, loops, rows_per_atom=1):
        self.nrow, self._dir, self.rhs = nrow, dir_, rhs
        self.families, self.groups, self.loops = families, groups, loops
        self.rows_per_atom = rows_per_atom              # > 1: vector atoms (`x[, b, 2] >= 1`), rows named atom[k]
        self.init_name, self.head = "", "["             # head: the row names up to the first loop ("make[", "hi[i=1,")
        self.n_terms = int(sum(f["count"] for f in families))

    @property
    def dir(self):
        return [self._dir] * self.nrow

    @property
    def names(self):
        return [self.init_name] * self.nrow

    @property
    def rownames(self):                                 # flatten_for_split + name_constraint (R/utils.R:66-94,154-165)
        from . import model
        parts = [[f"{l.name}={model._chr(v)}" for v in l.seq] for l in self.loops]
        atoms = [f"{self.head}{','.join(p)}]" for p in itertools.product(*parts)]
        if self.rows_per_atom == 1:
            return atoms
        return [f"{a}[{k}]" for a in atoms for k in range(1, self.rows_per_atom + 1)]


def build_block(low: LoweredFor) -> LoweredCon:
    """Families of one lowered block, rows numbered from 0 (easylp._csr places the block)."""
    for_loops = low.loops
    outer = low.loops + list(low.con.row_loops)         # a vector atom's own rows vary fastest
    nrow = int(np.prod([len(l.seq) for l in outer]))
    o_axes = {l.uid: i for i, l in enumerate(outer)}
    rhs = np.broadcast_to(_evaluate(low.con.rhs, o_axes, len(outer)), [len(l.seq) for l in outer]).reshape(-1).astype(float)
    if np.isnan(rhs).any():
        raise NotLowerable()
    # row stride of outer loop i: rows are numbered with the first loop outermost
    rstride, s = [0] * len(outer), 1
    for i in range(len(outer) - 1, -1, -1):
        rstride[i] = s
        s *= len(outer[i].seq)
    if len(low.con.groups) > 256:
        raise NotLowerable()                            # hundreds of separate addends per row: the eager path copes better
    families, groups, offset = [], [], 0
    outer_ext = [len(l.seq) for l in outer]
    for g in low.con.groups:
        # the loop nest of this group's families, slowest first.  A ragged `sum_for` index (one set per value of an
        # outer loop) is fused with that outer loop into ONE trailing loop over the (parent, value) pairs: rows and
        # columns come from tables over the pairs.  Outer loops may be reordered freely (rows are computed, not
        # counted); the order of the terms INSIDE a row — the sum_for grid order — is what the fold depends on.
        rag = [l for l in g.loops if getattr(l, "ragged", None)]
        expand, fused = {}, None
        if rag:
            if len(rag) != 1 or g.loops[-1] is not rag[0]:
                raise NotLowerable()
            fused = rag[0]
            parent, counts = fused.ragged
            if counts.size != len(parent.seq):
                raise NotLowerable()
            ppos = np.repeat(np.arange(len(parent.seq), dtype=np.int64), counts)
            expand = {parent.uid: ppos}
            if parent.uid in o_axes:                    # one set per value of a `for` index: rows come from a table
                nest = [l for l in outer if l.uid != parent.uid] + list(g.loops[:-1])
                row_stride = [rstride[o_axes[l.uid]] if l.uid in o_axes else 0 for l in nest] + [0]
                row_tabs = [None] * len(nest) + [(ppos * rstride[o_axes[parent.uid]]).astype(_I)]
            elif len(g.loops) >= 2 and g.loops[-2] is parent:   # ... of the fastest `sum_for` index: all inside one row
                nest = outer + list(g.loops[:-2])
                row_stride = rstride + [0] * (len(g.loops) - 2) + [0]
                row_tabs = [None] * (len(nest) + 1)
            else:
                raise NotLowerable()
            axes = {l.uid: i for i, l in enumerate(nest)}
            axes[parent.uid] = axes[fused.uid] = len(nest)
            ext = [len(l.seq) for l in nest] + [len(fused.seq)]
            nest_uids = [[l.uid] for l in nest] + [[parent.uid, fused.uid]]
        else:
            nest = outer + list(g.loops)
            axes = {l.uid: i for i, l in enumerate(nest)}
            ext = [len(l.seq) for l in nest]
            row_stride = rstride + [0] * len(g.loops)
            row_tabs = [None] * len(nest)
            nest_uids = [[l.uid] for l in nest]
        if len(ext) > MAX_LOOPS:
            raise NotLowerable()
        cells = int(np.prod(ext))
        posts = []
        for k in g.post:
            free = k.syms({})
            if any(u not in o_axes for u in free):
                raise NotLowerable()
            v = _evaluate(k, o_axes, len(outer))
            if not np.all(np.isfinite(v)):
                raise NotLowerable()
            posts.append(np.broadcast_to(v, outer_ext).reshape(-1).astype(float) if v.size > 1 else v.reshape(-1).astype(float))
        groups.append(posts)
        for ti, t in enumerate(g.terms):
            if any(u not in axes for u in t.tabs):
                raise NotLowerable()
            free = t.coef.syms({})
            if any(u not in axes for u in free):
                raise NotLowerable()
            v = _evaluate(t.coef, axes, len(ext), expand)
            if not np.all(np.isfinite(v)):
                raise NotLowerable()
            cstride, st = [0] * len(ext), 1
            for i in range(len(ext) - 1, -1, -1):
                if v.shape[i] > 1:
                    cstride[i] = st
                    st *= v.shape[i]
            col_tabs = []
            for uids in nest_uids:
                tab = None
                for u in uids:
                    if u in t.tabs:
                        tu = t.tabs[u][1].astype(np.int64)
                        if u in expand:
                            tu = tu[expand[u]]
                        tab = tu if tab is None else tab + tu
                col_tabs.append(None if tab is None else tab.astype(_I))
            families.append(dict(count=cells, out_offset=offset + ti, out_stride=len(g.terms), group=len(groups) - 1,
                                 extent=ext, row_stride=row_stride, row_tabs=row_tabs, col0=t.col0, col_tabs=col_tabs,
                                 coef=np.ascontiguousarray(v, dtype=float).reshape(-1), coef_stride=cstride))
        offset += cells * len(g.terms)
    per_atom = int(np.prod([len(l.seq) for l in low.con.row_loops])) if low.con.row_loops else 1
    return LoweredCon(nrow, low.con.op, rhs, families, groups, for_loops, per_atom)


# ---- packing for the C ABI ---------------------------------------------------------------------------
class TermFamily(C.Structure):
    _fields_ = [("count", C.c_int64), ("out_offset", C.c_int64), ("coef_tab", C.c_int64),
                ("col_tab", C.c_int64 * MAX_LOOPS), ("row_tab", C.c_int64 * MAX_LOOPS),
                ("coef_stride", C.c_int64 * MAX_LOOPS),
                ("extent", C.c_int32 * MAX_LOOPS), ("row_stride", C.c_int32 * MAX_LOOPS),
                ("out_stride", C.c_int32), ("group", C.c_int32), ("n_loops", C.c_int32), ("row0", C.c_int32),
                ("col0", C.c_int32), ("reserved", C.c_int32)]


class FoldGroup(C.Structure):
    _fields_ = [("mul_tab", C.c_int64 * MAX_GROUP_MUL), ("mul_per_row", C.c_int32 * MAX_GROUP_MUL),
                ("n_mul", C.c_int32), ("row0", C.c_int32)]


def pack(blocks_with_rows):
    """[(LoweredCon, first row)] -> (families array, itab, dtab, groups array); group 0 is the plain group."""
    fams, itab, dtab = [], [], []
    groups = [FoldGroup()]                              # group 0: explicit terms, no multipliers
    ni = nd = 0
    stream = 0
    for blk, row0 in blocks_with_rows:
        gids = []
        for posts in blk.groups:
            g = FoldGroup()
            g.n_mul, g.row0 = len(posts), row0
            for k, arr in enumerate(posts):
                g.mul_tab[k], g.mul_per_row[k] = nd, 1 if arr.size > 1 else 0
                dtab.append(arr)
                nd += arr.size
            gids.append(len(groups))
            groups.append(g)
        for f in blk.families:
            t = TermFamily()
            t.count, t.out_offset, t.out_stride = f["count"], stream + f["out_offset"], f["out_stride"]
            t.group, t.n_loops, t.row0, t.col0 = gids[f["group"]], len(f["extent"]), row0, f["col0"]
            t.coef_tab = nd
            dtab.append(f["coef"])
            nd += f["coef"].size
            for l in range(len(f["extent"])):
                t.extent[l], t.row_stride[l], t.coef_stride[l] = f["extent"][l], f["row_stride"][l], f["coef_stride"][l]
                for key, dst in (("col_tabs", t.col_tab), ("row_tabs", t.row_tab)):
                    tab = f[key][l]
                    if tab is None:
                        dst[l] = -1
                    else:
                        dst[l] = ni
                        itab.append(tab)
                        ni += tab.size
            fams.append(t)
        stream += blk.n_terms
    fam_arr = (TermFamily * max(len(fams), 1))(*fams)
    grp_arr = (FoldGroup * len(groups))(*groups)
    itab = np.concatenate(itab).astype(_I) if itab else np.zeros(0, _I)
    dtab = np.concatenate(dtab).astype(float) if dtab else np.zeros(0)
    return fam_arr, len(fams), itab, dtab, grp_arr, len(groups), stream
