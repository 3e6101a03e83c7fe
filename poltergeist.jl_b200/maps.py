"""Host-side mirror of the reference map constructors (src/maps/*.jl, src/poetry.jl).

Same names, argument meaning and error behaviour as the reference; what they produce
is a `pgb_map_desc` (include/poltergeist_b200.h) for the CUDA engine instead of a tree
of Julia closures.  Closures never cross the C ABI: a branch is either one of the
built-in parametric families (inverted in-kernel by FP64 Newton / closed form) or an
arbitrary Python callable that is pre-sampled here into Chebyshev coefficients of its
*inverse* branch and of |v'| and evaluated in-kernel by Clenshaw.

  MarkovMap / IntervalMap      src/maps/IntervalMap.jl:31-106
  modulomap / FwdOffset        src/maps/IntervalMap.jl:174-202,244-245
  CircleMap (Fwd/Rev)          src/maps/CircleMap.jl:5-92
  ComposedMarkovMap / compose  src/maps/ComposedMaps.jl:2-19
  perturb                      src/poetry.jl:37-46
  inv (single branch maps)     src/maps/conjugacy.jl
  lanford / tupling / doubling src/maps/examples.jl:4-37
"""
import ctypes as C
import math

import numpy as np

from . import _abi as A

Forward, Reverse = A.FORWARD, A.REVERSE
EPS = float(np.finfo(np.float64).eps)


def jl_eps(x):
    """Base.eps(::Float64)."""
    x = abs(float(x))
    if x == 0.0:
        return 5e-324
    return float(np.nextafter(x, np.inf) - x)


# ---- domains -------------------------------------------------------------------
class Interval:
    periodic = False

    def __init__(self, a, b):
        self.a, self.b = float(a), float(b)

    @property
    def arclength(self):
        return self.b - self.a

    def __contains__(self, x):
        return self.a <= x <= self.b

    def __eq__(self, o):
        return type(o) is type(self) and o.a == self.a and o.b == self.b

    def __hash__(self):
        return hash((type(self).__name__, self.a, self.b))

    def __repr__(self):
        return f"{self.a}..{self.b}"


class PeriodicSegment(Interval):
    periodic = True

    def __repr__(self):
        return f"PeriodicSegment({self.a},{self.b})"


def _dom(d, periodic=None):
    if isinstance(d, Interval):
        return d
    a, b = d
    return PeriodicSegment(a, b) if periodic else Interval(a, b)


def coveringinterval(ds):
    """src/general.jl:63-64"""
    ds = [_dom(d) for d in ds]
    return Interval(min(d.a for d in ds), max(d.b for d in ds))


def approx_in(x, d, atol=None):
    """src/general.jl:52-55"""
    atol = 50 * jl_eps(d.arclength) if atol is None else atol
    return d.a - atol <= x <= d.b + atol


def interval_mod(x, a, b):
    """src/general.jl:47"""
    return (b - a) * np.mod((x - a) / (b - a), 1.0) + a


# ---- forward-mode dual numbers: FunctionDerivative, src/general.jl:177-180 -------
class Dual:
    __array_priority__ = 1000
    _rules = {
        np.sin: lambda x: (np.sin(x), np.cos(x)), np.cos: lambda x: (np.cos(x), -np.sin(x)),
        np.tan: lambda x: (np.tan(x), 1 / np.cos(x) ** 2),
        np.exp: lambda x: (np.exp(x), np.exp(x)), np.log: lambda x: (np.log(x), 1 / x),
        np.sqrt: lambda x: (np.sqrt(x), 0.5 / np.sqrt(x)),
        np.arcsin: lambda x: (np.arcsin(x), 1 / np.sqrt(1 - x * x)),
        np.arccos: lambda x: (np.arccos(x), -1 / np.sqrt(1 - x * x)),
        np.arctan: lambda x: (np.arctan(x), 1 / (1 + x * x)),
        np.sinh: lambda x: (np.sinh(x), np.cosh(x)), np.cosh: lambda x: (np.cosh(x), np.sinh(x)),
        np.tanh: lambda x: (np.tanh(x), 1 / np.cosh(x) ** 2),
    }

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def _lift(o):
        return o if isinstance(o, Dual) else Dual(o, 0.0)

    def __add__(self, o):
        o = Dual._lift(o)
        return Dual(self.v + o.v, self.d + o.d)

    __radd__ = __add__

    def __sub__(self, o):
        o = Dual._lift(o)
        return Dual(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        return Dual._lift(o) - self

    def __mul__(self, o):
        o = Dual._lift(o)
        return Dual(self.v * o.v, self.d * o.v + self.v * o.d)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Dual._lift(o)
        return Dual(self.v / o.v, (self.d * o.v - self.v * o.d) / (o.v * o.v))

    def __rtruediv__(self, o):
        return Dual._lift(o) / self

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __pow__(self, p):
        if isinstance(p, Dual):
            raise TypeError("Dual ** Dual not supported")
        return Dual(self.v ** p, p * self.v ** (p - 1) * self.d)

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__":
            return NotImplemented
        if ufunc in Dual._rules and len(inputs) == 1:
            f, df = Dual._rules[ufunc](inputs[0].v)
            return Dual(f, df * inputs[0].d)
        table = {np.add: "__add__", np.subtract: "__sub__", np.multiply: "__mul__", np.true_divide: "__truediv__"}
        if ufunc in table and len(inputs) == 2:
            a, b = Dual._lift(inputs[0]), inputs[1]
            return getattr(a, table[ufunc])(b)
        if ufunc is np.negative:
            return -inputs[0]
        if ufunc is np.power:
            return inputs[0] ** inputs[1]
        return NotImplemented


def autodiff(f):
    """autodiff(f, d): derivative by dual numbers (src/maps/ExpandingBranch.jl:88-102)."""
    def df(x):
        try:
            out = f(Dual(x, np.ones_like(x) if hasattr(x, "shape") else 1.0))
        except TypeError as e:  # reference: "your function must accept DualNumbers"
            raise TypeError("To use automatic differentiation, your function must accept Dual numbers "
                            "(use numpy ufuncs, not math.*)") from e
        if not isinstance(out, Dual):
            return np.zeros_like(x) if hasattr(x, "shape") else 0.0
        return out.d
    return df


# ---- branch function families ----------------------------------------------------
class Family:
    """A branch function the kernels know in closed form."""
    family = None

    def par(self):
        raise NotImplementedError

    def __call__(self, x):
        raise NotImplementedError

    def deriv(self, x):
        raise NotImplementedError

    def with_offset(self, off):
        raise NotImplementedError


class PolySine(Family):
    """f(x) = p0 + p1 x + p2 x^2 + p3 x^3 + A sin(w x + phi) - offset"""
    family = A.FAM_POLYSINE

    def __init__(self, p0=0.0, p1=0.0, p2=0.0, p3=0.0, A=0.0, w=2 * math.pi, phi=0.0, offset=0.0):
        self.p = [float(p0), float(p1), float(p2), float(p3), float(A), float(w), float(phi), float(offset)]

    def par(self):
        return list(self.p)

    def __call__(self, x):
        p = self.p
        s = np.sin(p[5] * x + p[6]) if p[4] != 0 else 0.0
        return ((p[3] * x + p[2]) * x + p[1]) * x + p[0] + p[4] * s - p[7]

    def deriv(self, x):
        p = self.p
        c = np.cos(p[5] * x + p[6]) if p[4] != 0 else 0.0
        return (3.0 * p[3] * x + 2.0 * p[2]) * x + p[1] + p[4] * p[5] * c

    def with_offset(self, off):
        q = PolySine(*self.p)
        q.p[7] = self.p[7] + float(off)
        return q


class SqrtFam(Family):
    """f(x) = q0 + q1 sqrt(q2 + q3 x)  (Lanford inverse branches, examples.jl:4-8)"""
    family = A.FAM_SQRT

    def __init__(self, q0, q1, q2, q3):
        self.q = [float(q0), float(q1), float(q2), float(q3)]

    def par(self):
        return self.q + [0.0] * 4

    def __call__(self, x):
        q = self.q
        return q[0] + q[1] * np.sqrt(q[2] + q[3] * x)

    def deriv(self, x):
        q = self.q
        return q[1] * q[3] / (2.0 * np.sqrt(q[2] + q[3] * x))


class SinAsin(Family):
    """f(x) = sin(c0 + c1 asin x)  (test/runtests.jl:111)"""
    family = A.FAM_SINASIN

    def __init__(self, c0, c1):
        self.c = [float(c0), float(c1)]

    def par(self):
        return self.c + [0.0] * 6

    def __call__(self, x):
        return np.sin(self.c[0] + self.c[1] * np.arcsin(x))

    def deriv(self, x):
        return self.c[1] * np.cos(self.c[0] + self.c[1] * np.arcsin(x)) / np.sqrt(1.0 - x * x)


class Mobius(Family):
    """f(x) = (p0 + p1 x)/(p2 + p3 x)"""
    family = A.FAM_MOBIUS

    def __init__(self, p0, p1, p2, p3):
        self.p = [float(p0), float(p1), float(p2), float(p3)]

    def par(self):
        return self.p + [0.0] * 4

    def __call__(self, x):
        p = self.p
        return (p[0] + p[1] * x) / (p[2] + p[3] * x)

    def deriv(self, x):
        p = self.p
        return (p[1] * p[2] - p[0] * p[3]) / (p[2] + p[3] * x) ** 2


class ChebSeries(Family):
    """Chebyshev series on [a,b] for the value and (separately) for the derivative."""
    family = A.FAM_CHEB

    def __init__(self, a, b, coef, dcoef):
        self.a, self.b = float(a), float(b)
        self.coef = np.ascontiguousarray(coef, dtype=np.float64)
        self.dcoef = np.ascontiguousarray(dcoef, dtype=np.float64)

    def par(self):
        return [self.a, self.b] + [0.0] * 6

    def _t(self, x):
        return (2.0 * x - self.a - self.b) / (self.b - self.a)

    def __call__(self, x):
        return np.polynomial.chebyshev.chebval(self._t(x), self.coef)

    def deriv(self, x):
        return np.polynomial.chebyshev.chebval(self._t(x), self.dcoef)


class Closure:
    """An arbitrary host callable (+ optional derivative); pre-sampled, never shipped."""

    def __init__(self, f, df=None, offset=0.0):
        self.f, self.offset = f, float(offset)
        self.df = df if df is not None else autodiff(f)

    def __call__(self, x):
        return self.f(x) - self.offset

    def deriv(self, x):
        return self.df(x)

    def with_offset(self, off):
        return Closure(self.f, self.df, self.offset + off)


def _asfun(f, diff=None):
    if isinstance(f, (Family, Closure)):
        return f
    if callable(f):
        return Closure(f, diff)
    raise TypeError(f"cannot use {f!r} as a branch function")


# ---- Newton on the host (setup-time only: breakpoints, sampling) ------------------
def interval_newton(f, df, y, da, db, x=None, tol=None):
    """src/general.jl:117-131, scalar; raises like the reference when not converged."""
    x = (da + db) / 2 if x is None else float(x)
    tol = 40 * jl_eps(max(abs(da), abs(db))) if tol is None else tol
    rem = float(f(x)) - y
    for i in range(1, 26):
        x -= rem / float(df(x)) + jl_eps(x) * ((i % 11) - 5) / 4
        x = min(max(x, da), db)
        rem = float(f(x)) - y
        if abs(rem) < tol:
            break
    if abs(rem) > tol:
        raise ArithmeticError(f"Newton: failure to converge: y = {y}, x estimate = {x}, rem = {rem}")
    return x


def _solve_monotone(g, dg, y, da, db):
    """vectorised safeguarded Newton for g(v)=y on [da,db] in extended precision."""
    LD = np.longdouble
    y = np.asarray(y, dtype=LD)
    lo = np.full(y.shape, LD(da))
    hi = np.full(y.shape, LD(db))
    inc = g(LD(db)) > g(LD(da))
    x = (lo + hi) / 2
    for _ in range(200):
        r = np.asarray(g(x), dtype=LD) - y
        up = (r > 0) == inc
        hi = np.where(up, x, hi)
        lo = np.where(up, lo, x)
        d = np.asarray(dg(x), dtype=LD)
        with np.errstate(divide="ignore", invalid="ignore"):
            xn = x - r / d
        bad = ~((xn > lo) & (xn < hi))
        xn = np.where(bad, (lo + hi) / 2, xn)
        xn = np.where(r == 0, x, xn)
        done = np.abs(xn - x) <= 4 * np.finfo(LD).eps * np.maximum(np.abs(x), LD(1e-300))
        x = xn
        if bool(np.all(done)):
            break
    return x


def _cheb_nodes(n):
    j = np.arange(n, dtype=np.longdouble)
    return np.cos(np.longdouble(np.pi) * 0 + (np.pi * (2 * np.arange(n) + 1) / (2 * n)).astype(np.longdouble)), j


def _cheb_coeffs(vals):
    """first-kind-node values (node j = cos(pi (j+1/2)/n)) -> Chebyshev coefficients."""
    from scipy.fft import dct
    n = len(vals)
    c = dct(np.asarray(vals, dtype=np.float64), type=2) / n
    c[0] *= 0.5
    return c


class PresampleError(ValueError):
    """a closure branch whose inverse (or derivative) is not analytic up to the ends of the branch: its Chebyshev series does
    not converge to round-off, and the operator built from it would be silently inaccurate.  The reference's per-node Newton
    (ExpandingBranch.jl:34-37) stays exact there; use a parametric family (in-kernel Newton) for such a branch."""


PRESAMPLE_TAIL_LIMIT = 1e-10   # relative tail a pre-sampled series may be left with before it is refused outright


def _presample_not_converged(what, n, cv, cw):
    import warnings
    lvl = max(np.max(np.abs(c[-max(3, n // 8):])) / max(np.max(np.abs(c)), 1e-300) for c in (cv, cw))
    msg = (f"pre-sampled {what}: the Chebyshev series has not converged at {n} nodes (relative tail {float(lvl):.1e}); the branch or its "
           "derivative is not analytic up to the ends of its domain")
    if lvl > PRESAMPLE_TAIL_LIMIT:
        raise PresampleError(msg + " -- refused; give the branch as a parametric family instead")
    warnings.warn(msg + f" -- entries of the operator are good to ~{float(lvl):.0e} only", RuntimeWarning, stacklevel=3)


def sample_inverse(g, dg, da, db, ya, yb, max_n=1 << 14):
    """Chebyshev coefficients on [ya,yb] of v = g^{-1} and of v' = 1/g'(v)."""
    n = 17
    while True:
        t = np.cos(np.pi * (2 * np.arange(n) + 1) / (2.0 * n))
        y = (np.longdouble(ya) + np.longdouble(yb)) / 2 + (np.longdouble(yb) - np.longdouble(ya)) / 2 * t.astype(np.longdouble)
        v = _solve_monotone(g, dg, y, da, db)
        w = 1 / np.asarray(dg(v), dtype=np.longdouble)
        cv, cw = _cheb_coeffs(v), _cheb_coeffs(w)
        ok = True
        for c in (cv, cw):
            tail = np.max(np.abs(c[-max(3, n // 8):]))
            ok &= tail <= 4 * EPS * max(np.max(np.abs(c)), 1e-300)
        if ok:
            break
        if n >= max_n:
            _presample_not_converged("inverse branch", n, cv, cw)
            break
        n = 2 * (n - 1) + 1
    def chop(c):
        thr = 0.25 * EPS * np.max(np.abs(c))
        k = len(c)
        while k > 1 and abs(c[k - 1]) <= thr:
            k -= 1
        return c[:k].copy()
    return chop(cv), chop(cw)


# ---- branches ---------------------------------------------------------------------
class Branch:
    """Fwd/RevExpandingBranch (src/maps/ExpandingBranch.jl:13-66)."""

    def __init__(self, fn, dom, ran, dir=Forward, mode=A.MODE_INTERVAL, shift=0.0, fa=0.0, fb=0.0):
        self.fn, self.dom, self.ran, self.dir = fn, _dom(dom), _dom(ran), dir
        self.mode, self.shift, self.fa, self.fb = mode, float(shift), float(fa), float(fb)

    def _emit(self, coefs):
        """-> pgb_branch; closures are sampled into FAM_CHEB here."""
        fn, d = self.fn, self.dir
        if isinstance(fn, Closure):
            if self.mode == A.MODE_INTERVAL and d == Forward:
                cv, cw = sample_inverse(fn, fn.deriv, self.dom.a, self.dom.b, self.ran.a, self.ran.b)
                fn, d = ChebSeries(self.ran.a, self.ran.b, cv, cw), Reverse
            elif self.mode == A.MODE_FWD_CIRCLE and d == Forward:
                lo, hi = min(self.fa, self.fb), max(self.fa, self.fb)
                cv, cw = sample_inverse(fn, fn.deriv, self.dom.a, self.dom.b, lo, hi)
                fn, d = ChebSeries(lo, hi, cv, cw), Reverse
            else:  # the closure already is the inverse: sample it directly
                lo = self.ran.a + (self.shift if self.mode == A.MODE_REV_CIRCLE else 0.0)
                hi = self.ran.b + (self.shift if self.mode == A.MODE_REV_CIRCLE else 0.0)
                fn = _sample_direct(fn, lo, hi)
        b = A.pgb_branch()
        b.family, b.dir, b.mode = fn.family, d, self.mode
        b.dom_a, b.dom_b, b.ran_a, b.ran_b = self.dom.a, self.dom.b, self.ran.a, self.ran.b
        b.shift, b.fa, b.fb = self.shift, self.fa, self.fb
        for i, p in enumerate(fn.par()):
            b.par[i] = p
        if isinstance(fn, ChebSeries):
            b.coef_off, b.ncoef = len(coefs), len(fn.coef)
            coefs.extend(fn.coef.tolist())
            b.dcoef_off, b.ndcoef = len(coefs), len(fn.dcoef)
            coefs.extend(fn.dcoef.tolist())
        return b


def _sample_direct(fn, lo, hi, max_n=1 << 14):
    n = 17
    while True:
        t = np.cos(np.pi * (2 * np.arange(n) + 1) / (2.0 * n))
        y = (lo + hi) / 2 + (hi - lo) / 2 * t
        cv, cw = _cheb_coeffs(fn(y)), _cheb_coeffs(fn.deriv(y))
        ok = all(np.max(np.abs(c[-max(3, n // 8):])) <= 4 * EPS * max(np.max(np.abs(c)), 1e-300) for c in (cv, cw))
        if not ok and n >= max_n:
            _presample_not_converged("branch", n, cv, cw)
        if ok or n >= max_n:
            return ChebSeries(lo, hi, cv, cw)
        n = 2 * (n - 1) + 1


# ---- maps -------------------------------------------------------------------------
class AbstractMarkovMap:
    periodic = False

    def stages(self):
        return [self]

    def compose(self, other):
        return ComposedMarkovMap(self, other)

    __matmul__ = compose

    @property
    def default_space(self):
        return A.FOURIER if self.domain.periodic else A.CHEBYSHEV

    def to_desc(self, domain_space=None, range_space=None):
        """pgb_map_desc + the buffers it points into (kept alive on the returned object)."""
        stages = self.stages()
        coefs, branches, st = [], [], []
        for s in stages:
            bl = s.branchlist()
            st.append((len(bl), len(branches)))
            branches.extend(b._emit(coefs) for b in bl)
        d = A.pgb_map_desc()
        d.nstage, d.nbranch, d.ncoef = len(stages), len(branches), len(coefs)
        dom, ran = stages[-1].domain, stages[0].rangedomain
        d.domain_space = (A.FOURIER if dom.periodic else A.CHEBYSHEV) if domain_space is None else domain_space
        d.range_space = (A.FOURIER if ran.periodic else A.CHEBYSHEV) if range_space is None else range_space
        d.dom_a, d.dom_b, d.ran_a, d.ran_b = dom.a, dom.b, ran.a, ran.b
        d._stages = (A.pgb_stage * len(st))(*[A.pgb_stage(n, o) for n, o in st])
        d._branches = (A.pgb_branch * len(branches))(*branches)
        d._coefs = (C.c_double * max(1, len(coefs)))(*coefs)
        d.stages = C.cast(d._stages, C.POINTER(A.pgb_stage))
        d.branches = C.cast(d._branches, C.POINTER(A.pgb_branch))
        d.coefs = C.cast(d._coefs, C.POINTER(C.c_double))
        return d


class MarkovMap(AbstractMarkovMap):
    """MarkovMap(fs, ds, ran=coveringinterval(ds); dir=Forward, diff=autodiff) -- full branches.
    src/maps/IntervalMap.jl:31-71"""
    markov = True

    def __init__(self, fs, ds, ran=None, dir=Forward, diff=None):
        if len(fs) != len(ds):
            raise AssertionError("length(fs) == length(ds)")
        ds = [_dom(d) for d in ds]
        self.domain = coveringinterval(ds)
        self.rangedomain = _dom(ran) if ran is not None else coveringinterval(ds)
        diff = diff if diff is not None else [None] * len(fs)
        self.branches = []
        for f, d, df in zip(fs, ds, diff):
            r = self.rangedomain if self.markov else self._branch_range(_asfun(f, df), d)
            self.branches.append(Branch(_asfun(f, df), d, r, dir))
        for b in self.branches:  # @assert issubset(b.domain, dom); b.rangedomain == ran
            assert b.dom.a >= self.domain.a and b.dom.b <= self.domain.b

    def branchlist(self):
        return self.branches

    def nbranches(self):
        return len(self.branches)

    def __repr__(self):
        return f"{type(self).__name__} {self.domain} → {self.rangedomain} with {self.nbranches()} branches"


class IntervalMap(MarkovMap):
    """IntervalMap: branch ranges are the images of the branch domains (IntervalMap.jl:76-106)."""
    markov = False

    def __init__(self, fs, ds, ran=None, dir=Forward, diff=None):
        if dir != Forward:
            raise NotImplementedError("IntervalMap with dir=Reverse needs explicit branch ranges")
        if ran is None:
            ran = coveringinterval([self._branch_range(_asfun(f), _dom(d)) for f, d in zip(fs, ds)])
        super().__init__(fs, ds, ran, dir, diff)
        for b in self.branches:
            assert approx_in(b.ran.a, self.rangedomain) and approx_in(b.ran.b, self.rangedomain)

    @staticmethod
    def _branch_range(f, d):
        """mapinterval(f, d), src/general.jl:65"""
        ya, yb = float(f(d.a)), float(f(d.b))
        return Interval(min(ya, yb), max(ya, yb))


class FwdCircleMap(AbstractMarkovMap):
    """src/maps/CircleMap.jl:5-37"""

    def __init__(self, f, dom, ran=None, diff=None):
        self.f = _asfun(f, diff)
        self.domain = _dom(dom, True)
        self.rangedomain = _dom(ran, True) if ran is not None else self.domain
        self.fa, self.fb = float(self.f(self.domain.a)), float(self.f(self.domain.b))
        cover_est = (self.fb - self.fa) / self.rangedomain.arclength
        ci = int(round(cover_est))
        if not math.isclose(cover_est, ci, rel_tol=math.sqrt(EPS)) or ci == 0:
            raise ValueError("Circle map lift does not have integer covering number.")
        self.cover = abs(ci)

    def branchlist(self):
        L = self.rangedomain.arclength
        return [Branch(self.f, self.domain, self.rangedomain, Forward, A.MODE_FWD_CIRCLE, i * L, self.fa, self.fb)
                for i in range(1, self.cover + 1)]

    def nbranches(self):
        return self.cover


class RevCircleMap(AbstractMarkovMap):
    """src/maps/CircleMap.jl:40-83"""

    def __init__(self, v, dom, ran=None, diff=None, maxcover=10000):
        self.v = _asfun(v, diff)
        self.domain = _dom(dom, True)
        self.rangedomain = _dom(ran, True) if ran is not None else self.domain
        ra, dr, dd = self.rangedomain.a, self.rangedomain.arclength, self.domain.arclength
        va = float(self.v(ra))
        cover, vr = 1, float(self.v(ra + dr)) - va
        while abs(vr) <= dd and cover < maxcover:  # CircleMap.jl:50-64
            vr = float(self.v(ra + cover * dr)) - va
            if math.isclose(abs(vr), dd, rel_tol=math.sqrt(EPS)):
                break
            if abs(vr) > dd:
                raise ValueError("Inverse lift doesn't appear to have an inverse")
            cover += 1
        if cover == maxcover:
            raise ValueError(f"Can't get to the end of the inverse lift after {maxcover} steps")
        self.cover = cover

    def branchlist(self):
        L = self.rangedomain.arclength
        return [Branch(self.v, self.domain, self.rangedomain, Reverse, A.MODE_REV_CIRCLE, i * L)
                for i in range(1, self.cover + 1)]

    def nbranches(self):
        return self.cover


def CircleMap(f, d, r=None, dir=Forward, diff=None):
    """src/maps/CircleMap.jl:91-92"""
    return FwdCircleMap(f, d, r, diff) if dir == Forward else RevCircleMap(f, d, r, diff)


class ComposedMarkovMap(AbstractMarkovMap):
    """f o g o ... (src/maps/ComposedMaps.jl:2-19): maps[0] is applied last."""

    def __init__(self, *maps):
        mps = []
        for m in maps:
            mps.extend(m.stages())
        for i in range(len(mps) - 1):
            if not (mps[i + 1].rangedomain == mps[i].domain):
                raise AssertionError("rangedomain(maps[i+1]) == domain(maps[i])")
        self.maps = tuple(mps)
        self.domain, self.rangedomain = mps[-1].domain, mps[0].rangedomain

    def stages(self):
        return list(self.maps)

    def nbranches(self):
        return int(np.prod([m.nbranches() for m in self.maps]))


def compose(*maps):
    return ComposedMarkovMap(*maps)


def forwardmodulomap(f, dom, ran=None, diff=None):
    """src/maps/IntervalMap.jl:180-202: branches of f mod ran, breakpoints by Newton."""
    f = _asfun(f, diff)
    domd = _dom(dom)
    randm = _dom(ran) if ran is not None else domd
    fa, fb = float(f(domd.a)), float(f(domd.b))
    L = randm.arclength
    nb = int(round((fb - fa) / L))
    if nb == 0:
        raise AssertionError("nb != 0")
    sg, NB = (1 if nb > 0 else -1), abs(nb)
    bp = [domd.a]
    for i in range(1, NB):
        bp.append(interval_newton(f, f.deriv, fa + sg * i * L, domd.a, domd.b,
                                  tol=40 * jl_eps(max(domd.a, domd.b))))
    bp.append(domd.b)
    fs = [f.with_offset((i - (1 if sg == 1 else 0)) * sg * L) for i in range(1, NB + 1)]
    ds = [Interval(bp[i], bp[i + 1]) for i in range(NB)]
    return MarkovMap(fs, ds, randm)


def modulomap(f, d, r=None, dir=Forward, diff=None):
    """src/maps/IntervalMap.jl:244-245"""
    if dir != Forward:
        raise NotImplementedError("not implemented: reversemodulomap")  # IntervalMap.jl:211
    return forwardmodulomap(f, d, r, diff)


def perturb(m_or_d, X, eps):
    """perturb(d, X, eps): x -> x + eps X(x) on d;  perturb(m, X, eps) = perturb(rangedomain(m),X,eps) o m
    (src/poetry.jl:37-46)."""
    if isinstance(m_or_d, AbstractMarkovMap):
        return ComposedMarkovMap(perturb(m_or_d.rangedomain, X, eps), m_or_d)
    d = _dom(m_or_d)
    if isinstance(X, PolySine) and all(X.p[i] == 0 for i in (2, 3)) :
        f = PolySine(eps * X.p[0], 1.0 + eps * X.p[1], 0, 0, eps * X.p[4], X.p[5], X.p[6], eps * X.p[7])
    else:
        Xf = X
        f = Closure(lambda x: x + eps * Xf(x))
    if d.periodic:
        return FwdCircleMap(f, d)
    return MarkovMap([f], [d], d)


def inv(m):
    """inverse of a single-branch map: swap Fwd <-> Rev (src/maps/conjugacy.jl)."""
    if isinstance(m, ComposedMarkovMap):
        return ComposedMarkovMap(*[inv(s) for s in reversed(m.maps)])
    if not isinstance(m, MarkovMap) or m.nbranches() != 1:
        raise NotImplementedError("inv is defined for single-branch interval maps")
    b = m.branches[0]
    out = MarkovMap.__new__(MarkovMap)
    out.domain, out.rangedomain = m.rangedomain, m.domain
    out.branches = [Branch(b.fn, b.ran, b.dom, Reverse if b.dir == Forward else Forward)]
    return out


# ---- examples (src/maps/examples.jl) --------------------------------------------------
sinpi = PolySine(A=1.0, w=math.pi)


def lanford():
    """lanford(): closed-form inverse branches v=(5-sqrt(n-8x))/2, n=25,17 (examples.jl:4-17)."""
    brk = (5.0 - math.sqrt(25.0 - 8.0)) / 2.0
    return MarkovMap([SqrtFam(2.5, -0.5, 25.0, -8.0), SqrtFam(2.5, -0.5, 17.0, -8.0)],
                     [Interval(0.0, brk), Interval(brk, 1.0)], dir=Reverse)


def tupling(k, d=None):
    """examples.jl:27-28"""
    d = _dom(d) if d is not None else Interval(0.0, 1.0)
    if d.periodic:
        return RevCircleMap(PolySine(p0=d.a - d.a / k, p1=1.0 / k), d, d)
    return modulomap(PolySine(p0=d.a - k * d.a, p1=float(k)), d)


def doubling(d=None):
    return tupling(2, d)
