"""oracle/dualnum.py -- TEST INFRASTRUCTURE, not product code.

Arrays of the reference's forward-mode dual numbers, Ibis::Dual<double>
(/root/reference/src/util/util/dual.h:11-346), for oracle/restatement.py: a pair of float64 arrays (re, du)
with the reference's operator set and operation ORDER -- `dual / dual` is not `dual / real`, `real - dual`
negates the dual part, comparisons look at the real part only, min / max keep the SECOND argument on a tie
(dual.h:269-298, unlike Kokkos::min / max for doubles), abs negates both parts, copysign flips the dual part by a
threshold test, pow carries the `power.dual * real * log(base)` term.

numpy dispatches its functions and ufuncs here (__array_ufunc__ / __array_function__), so that the restatement's
formulas run unchanged on real arrays and on Dual arrays.  A plain float / ndarray operand stands for a `double`
of the reference (literal or Ibis::real); where the reference's operand is a T (= Dual with zero dual part:
geometry, gas constants) and the operation is a DIVISION by it, the restatement promotes it with `as_T`.
"""
import numpy as np

HANDLED = {}


def implements(np_function):
    def decorator(func):
        HANDLED[np_function] = func
        return func
    return decorator


def _parts(x):
    return (x.re, x.du) if isinstance(x, Dual) else (np.asarray(x, dtype=np.float64), None)


class Dual:
    __slots__ = ("re", "du")
    __array_priority__ = 1000

    def __init__(self, re, du=None):
        self.re = np.array(re, dtype=np.float64)
        self.du = np.zeros_like(self.re) if du is None else np.array(du, dtype=np.float64)

    # ---- container behaviour
    @property
    def shape(self):
        return self.re.shape

    def __len__(self):
        return len(self.re)

    def copy(self):
        return Dual(self.re.copy(), self.du.copy())

    def __getitem__(self, idx):
        return Dual(self.re[idx], self.du[idx])

    def __setitem__(self, idx, value):
        if isinstance(value, Dual):
            self.re[idx], self.du[idx] = value.re, value.du
        else:                                                    # a double: Dual(x) has a zero dual part (dual.h:18)
            self.re[idx] = value
            self.du[idx] = 0.0

    # ---- comparisons: real parts only (dual.h:44-74)
    def _cmp(self, other, op):
        return op(self.re, other.re if isinstance(other, Dual) else other)

    def __lt__(self, o): return self._cmp(o, np.less)
    def __le__(self, o): return self._cmp(o, np.less_equal)
    def __gt__(self, o): return self._cmp(o, np.greater)
    def __ge__(self, o): return self._cmp(o, np.greater_equal)
    def __eq__(self, o): return self._cmp(o, np.equal)
    def __ne__(self, o): return self._cmp(o, np.not_equal)
    __hash__ = None

    # ---- arithmetic (dual.h:75-199)
    def __neg__(self):
        return Dual(-self.re, -self.du)

    def __add__(self, o):
        return Dual(self.re + o.re, self.du + o.du) if isinstance(o, Dual) else Dual(self.re + o, self.du)

    def __radd__(self, o):
        return Dual(o + self.re, self.du)

    def __sub__(self, o):
        return Dual(self.re - o.re, self.du - o.du) if isinstance(o, Dual) else Dual(self.re - o, self.du)

    def __rsub__(self, o):
        return Dual(o - self.re, -self.du)

    def __isub__(self, o):                                       # in place: the caller's other names see it
        r = self - o
        self.re[...], self.du[...] = r.re, r.du
        return self

    def __mul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.re * o.re, self.re * o.du + self.du * o.re)
        return Dual(self.re * o, self.du * o)

    def __rmul__(self, o):
        return Dual(self.re * o, self.du * o)

    def __truediv__(self, o):
        if isinstance(o, Dual):
            return Dual(self.re / o.re, (self.du * o.re - self.re * o.du) / (o.re * o.re))
        return Dual(self.re / o, self.du / o)

    def __rtruediv__(self, o):
        return Dual(o / self.re, -self.du * o / (self.re * self.re))

    # ---- numpy protocol
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        a = inputs[0]
        b = inputs[1] if len(inputs) > 1 else None
        if ufunc is np.add:
            return a + b if isinstance(a, Dual) else b.__radd__(a)
        if ufunc is np.subtract:
            return a - b if isinstance(a, Dual) else b.__rsub__(a)
        if ufunc is np.multiply:
            return a * b if isinstance(a, Dual) else b.__rmul__(a)
        if ufunc in (np.divide, np.true_divide):
            return a / b if isinstance(a, Dual) else b.__rtruediv__(a)
        if ufunc is np.negative:
            return -a
        if ufunc is np.absolute:                                 # dual.h:215: (real < 0) ? -(*this) : *this
            neg = a.re < 0.0
            return Dual(np.where(neg, -a.re, a.re), np.where(neg, -a.du, a.du))
        if ufunc is np.sqrt:                                     # dual.h:225-230
            r = np.sqrt(a.re)
            return Dual(r, a.du / (2.0 * r))
        if ufunc is np.floor:                                    # dual.h:300-303
            return Dual(np.floor(a.re), np.floor(a.du))
        if ufunc is np.copysign:                                 # dual.h:310-314
            m = a if isinstance(a, Dual) else Dual(np.broadcast_to(np.float64(a), b.re.shape))
            r = np.copysign(m.re, b.re)
            return Dual(r, np.where((r - m.re) < 1e-15, m.du, -m.du))
        if ufunc is np.power:                                    # dual.h:237-244 with power = Dual(p, 0)
            p = float(b)
            real = np.power(a.re, p)
            return Dual(real, a.du * p * np.power(a.re, p - 1.0) + 0.0 * real * np.log(a.re))
        if ufunc in (np.less, np.less_equal, np.greater, np.greater_equal, np.equal, np.not_equal):
            ar = a.re if isinstance(a, Dual) else a
            br = b.re if isinstance(b, Dual) else b
            return ufunc(ar, br)
        return NotImplemented

    def __array_function__(self, func, types, args, kwargs):
        if func not in HANDLED:
            return NotImplemented
        return HANDLED[func](*args, **kwargs)


@implements(np.where)
def _where(cond, x, y):
    xr, xd = _parts(x)
    yr, yd = _parts(y)
    xd = np.zeros_like(xr) if xd is None else xd
    yd = np.zeros_like(yr) if yd is None else yd
    return Dual(np.where(cond, xr, yr), np.where(cond, xd, yd))


@implements(np.stack)
def _stack(arrays, axis=0):
    parts = [_parts(a) for a in arrays]
    return Dual(np.stack([p[0] for p in parts], axis=axis),
                np.stack([np.zeros_like(p[0]) if p[1] is None else p[1] for p in parts], axis=axis))


@implements(np.zeros_like)
def _zeros_like(a, **kw):
    return np.zeros_like(a.re)


@implements(np.ones_like)
def _ones_like(a, **kw):
    return np.ones_like(a.re)


def as_T(x, like):
    """x as the reference's T when the computation runs on duals: Dual(x, 0) -- needed where T DIVIDES a dual
    (dual / dual, dual.h:169-175, is not dual / real)"""
    return Dual(x) if isinstance(like, Dual) and not isinstance(x, Dual) else x


def dmin(a, b):
    """Ibis::min on duals (dual.h:284-298): a.real < b.real ? a : b  -- also for one double argument"""
    ar = a.re if isinstance(a, Dual) else a
    br = b.re if isinstance(b, Dual) else b
    return np.where(ar < br, a, b)


def dmax(a, b):
    """Ibis::max on duals (dual.h:269-282): a.real > b.real ? a : b"""
    ar = a.re if isinstance(a, Dual) else a
    br = b.re if isinstance(b, Dual) else b
    return np.where(ar > br, a, b)
