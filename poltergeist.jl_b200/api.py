"""Host-side mirror of the reference's operator interface for the hot path: same names, argument
meaning and error behaviour as Poltergeist.jl, with every matrix entry, QR factor and solve coming
from libpoltergeist_b200.so (CUDA, sm_100a).  Nothing here computes a transfer-matrix entry on the
CPU and nothing imports the test oracle; the only host arithmetic is the O(N) `Fun` algebra the
reference also does on the host (products, derivatives, integrals of coefficient vectors).

  Fun                 ApproxFun.Fun on Chebyshev(a..b) / Fourier(PeriodicSegment(a,b)) -- coefficient vector
  Transfer(m)         cache(ConcreteTransfer(m))          src/spectral/Transfer.jl:12-40,183-223
  SolutionInv(L[,u])  qr(I - L + u/sum(u) * integral)     src/spectral/Solution.jl:4-38
  acim, linearresponse, correlationsum, birkhoffcov, birkhoffvar
                                                           src/spectral/statistics.jl:9-63
  eigvals(L|m, n)                                          src/poetry.jl:56-66
  transfer(m, fn)                                          src/spectral/Transfer.jl:75-83 (via L * Fun(fn))
  BatchSweep          the same path over many maps of one shape (perturb sweeps, README.md:57-66)
"""
import ctypes as C

import numpy as np

from . import _abi as A
from . import engine as E
from .maps import AbstractMarkovMap, Interval, PeriodicSegment

EPS = float(np.finfo(np.float64).eps)


# ----------------------------------------------------------------------------------------- Fun
def _points(space, a, b, n):
    if space == A.CHEBYSHEV:
        i = np.arange(n)
        return (a + b) / 2 + (b - a) / 2 * np.sin(np.pi * (n - 2 * i - 1) / (2.0 * n))
    return a + (b - a) * np.arange(n) / n


def _transform(space, vals):
    """ApproxFun.transform (SURVEY Appendix B conventions)."""
    from scipy.fft import dct, rfft
    vals = np.asarray(vals, dtype=np.float64)
    n = len(vals)
    if space == A.CHEBYSHEV:
        c = dct(vals, type=2) / n
        c[0] *= 0.5
        return c
    F = rfft(vals)
    c = np.zeros(n)
    c[0] = F[0].real / n
    m = np.arange(1, (n - 1) // 2 + 1)
    c[2 * m - 1] = -2.0 * F[m].imag / n
    c[2 * m] = 2.0 * F[m].real / n
    if n % 2 == 0 and n > 1:
        c[n - 1] = F[n // 2].real / n
    return c


def _chop(c, tol):
    k = len(c)
    while k > 1 and abs(c[k - 1]) <= tol:
        k -= 1
    return np.array(c[:k], dtype=np.float64)


class Fun:
    """Coefficient vector on Chebyshev(a..b) (space CHEBYSHEV) or Fourier(PeriodicSegment(a,b))."""

    def __init__(self, f, domain=None, space=None, n=None):
        if isinstance(f, Fun):
            self.space, self.a, self.b, self.coefficients = f.space, f.a, f.b, f.coefficients.copy()
            return
        d = domain if isinstance(domain, Interval) else (Interval(*domain) if domain is not None else Interval(-1.0, 1.0))
        self.a, self.b = d.a, d.b
        self.space = (A.FOURIER if d.periodic else A.CHEBYSHEV) if space is None else space
        if callable(f):
            self.coefficients = self._sample(f, n)
        else:
            self.coefficients = np.atleast_1d(np.asarray(f, dtype=np.float64)).copy()

    def _sample(self, f, n):
        if n is not None:
            return _transform(self.space, np.asarray(f(_points(self.space, self.a, self.b, n)), dtype=np.float64) * np.ones(n))
        n = 16
        while True:
            c = _transform(self.space, np.asarray(f(_points(self.space, self.a, self.b, n)), dtype=np.float64) * np.ones(n))
            mx = max(np.abs(c).max(), 1e-300)
            tail = np.abs(c[-max(3, n // 8):]).max()
            if tail <= 10 * EPS * mx or n >= (1 << 16):
                return _chop(c, 2 * EPS * mx)
            n *= 2

    @property
    def domain(self):
        return (PeriodicSegment if self.space == A.FOURIER else Interval)(self.a, self.b)

    def ncoefficients(self):
        return len(self.coefficients)

    def _like(self, c):
        out = Fun.__new__(Fun)
        out.space, out.a, out.b, out.coefficients = self.space, self.a, self.b, np.asarray(c, dtype=np.float64)
        return out

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        t = (self.a + self.b - 2.0 * x) / (self.a - self.b)
        c = self.coefficients
        if self.space == A.CHEBYSHEV:
            return np.polynomial.chebyshev.chebval(t, c)
        th = np.pi * (t + 1.0)
        s = np.full_like(th, c[0])
        for k in range(1, len(c)):
            m = (k + 1) // 2
            s = s + c[k] * (np.sin(m * th) if k & 1 else np.cos(m * th))
        return s

    def _pad(self, n):
        c = np.zeros(n)
        k = min(n, len(self.coefficients))
        c[:k] = self.coefficients[:k]
        return c

    def _same(self, o):
        if not (isinstance(o, Fun) and o.space == self.space and o.a == self.a and o.b == self.b):
            raise ValueError("Funs live on different spaces")

    def __add__(self, o):
        if not isinstance(o, Fun):
            c = self.coefficients.copy(); c[0] += float(o)
            return self._like(c)
        self._same(o)
        n = max(len(self.coefficients), len(o.coefficients))
        return self._like(self._pad(n) + o._pad(n))

    __radd__ = __add__

    def __neg__(self):
        return self._like(-self.coefficients)

    def __sub__(self, o):
        return self + (-o)

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if not isinstance(o, Fun):
            return self._like(self.coefficients * float(o))
        self._same(o)
        f, g = self.coefficients, o.coefficients
        if self.space == A.CHEBYSHEV:
            p = np.polynomial.chebyshev.chebmul(f, g)
        else:
            m = 2
            while m < len(f) + len(g) + 2:
                m *= 2
            x = _points(A.FOURIER, self.a, self.b, m)
            p = _transform(A.FOURIER, self(x) * o(x))
        mx = max(np.abs(p).max(), 1e-300)
        return self._like(_chop(p, EPS * mx))

    __rmul__ = __mul__

    def __truediv__(self, s):
        return self._like(self.coefficients / float(s))

    def derivative(self):
        """Fun' (used by linearresponse, statistics.jl:19)"""
        c = self.coefficients
        if self.space == A.CHEBYSHEV:
            d = np.polynomial.chebyshev.chebder(c) * (2.0 / (self.b - self.a)) if len(c) > 1 else np.zeros(1)
            return self._like(d)
        d = np.zeros(len(c) + (len(c) % 2 == 0))
        sc = 2 * np.pi / (self.b - self.a)
        for k in range(1, len(c)):
            m = (k + 1) // 2
            if k & 1:
                d[2 * m] += sc * m * c[k]
            else:
                d[2 * m - 1] -= sc * m * c[k]
        return self._like(d)

    def sum(self):
        """DefiniteIntegral"""
        return float(integral_functional(self.space, self.a, self.b, len(self.coefficients)) @ self.coefficients)


def integral_functional(space, a, b, n):
    w = np.zeros(n)
    if space == A.CHEBYSHEV:
        k = np.arange(0, n, 2)
        w[0::2] = (b - a) / 2 * 2.0 / (1.0 - k.astype(float) ** 2)
    else:
        w[0] = b - a
    return w


def uniform(domain_or_fun):
    """uniform(S): the constant 1/|d| (Solution.jl:17-22)"""
    d = domain_or_fun.domain if isinstance(domain_or_fun, Fun) else domain_or_fun
    u = Fun([1.0], d)
    return u / u.sum()


# ----------------------------------------------------------------------------------------- Transfer
class Transfer:
    """Transfer(m) = cache(ConcreteTransfer(m)): columns are produced on demand on the GPU with the
    reference's adaptive node count and stay resident in HBM (Transfer.jl:33,201-223)."""

    def __init__(self, m, domain_space=None, range_space=None, ctx=None, padding=False):
        """padding: Transfer(stuff...; padding) (Transfer.jl:33) -- ragged() pads every column to the running maximum
        colstop (transfer_getindex's `padding && pad!(cutcfc, K)`, Transfer.jl:172-173)"""
        self.padding = bool(padding)
        if isinstance(m, Transfer):
            m = m.m
        if not isinstance(m, AbstractMarkovMap):
            raise TypeError("Transfer expects a MarkovMap / CircleMap / composition")
        self.m = m
        self.dm = E.DeviceMap(m, ctx, domain_space, range_space)
        self.ctx = self.dm.ctx
        self.domainspace, self.rangespace = self.dm.desc.domain_space, self.dm.desc.range_space
        self._h = C.c_void_p()
        A.check(A.lib().pgb_transfer_create(self.ctx.h, self.dm.h, C.byref(self._h)), self.ctx.h)

    def set_speculation(self, enable):
        A.check(A.lib().pgb_transfer_set_speculation(self._h, 1 if enable else 0), self.ctx.h)
        return self

    # -- cache protocol
    def resizedata(self, n):
        A.check(A.lib().pgb_transfer_resize(self._h, int(n)), self.ctx.h)
        return self

    @property
    def datasize(self):
        return int(A.lib().pgb_transfer_ncols(self._h))

    def colstop(self, k):
        """colstop of 0-based column k (fills up to it when unknown; Transfer.jl:219-223)"""
        cs = C.c_int32(0)
        A.check(A.lib().pgb_transfer_colstop(self._h, int(k), C.byref(cs)), self.ctx.h)
        return cs.value

    def nodes_used(self, lo, hi):
        out = np.zeros(hi - lo, dtype=np.int32)
        A.check(A.lib().pgb_transfer_nodes_used(self._h, lo, hi, out.ctypes.data_as(C.POINTER(C.c_int32))), self.ctx.h)
        return out

    def ragged(self, lo, hi):
        """RaggedMatrix(data, cols, m) of columns [lo, hi): cols 1-based, cols[0] = 1 (Transfer.jl:99-101)"""
        nnz = C.c_int64(0)
        A.check(A.lib().pgb_transfer_ragged_size(self._h, lo, hi, C.byref(nnz)), self.ctx.h)
        data = np.empty(max(nnz.value, 1))
        cols = np.empty(hi - lo + 1, dtype=np.int64)
        m = C.c_int64(0)
        A.check(A.lib().pgb_transfer_get_ragged(self._h, lo, hi, A.dptr(data), len(data), cols.ctypes.data_as(C.POINTER(C.c_int64)),
                                                C.byref(m)), self.ctx.h)
        if not self.padding:
            return data[:nnz.value], cols, m.value
        # pad!(cutcfc, K): K = the largest colstop seen so far, over the earlier columns too (Transfer.jl:104,171-173)
        K = max([self.colstop(c) for c in range(lo)] + [0])
        out, pcols = [], [1]
        for j in range(hi - lo):
            col = data[cols[j] - 1:cols[j + 1] - 1]
            K = max(K, len(col))
            out.append(np.pad(col, (0, K - len(col))))
            pcols.append(pcols[-1] + K)
        return (np.concatenate(out) if out else np.empty(0)), np.array(pcols, dtype=np.int64), K

    def dense(self, rows, lo, hi=None):
        """L[0:rows, lo:hi] on the host"""
        if hi is None:
            lo, hi = 0, lo
        out = np.zeros((hi - lo, max(rows, 1)))
        A.check(A.lib().pgb_transfer_get_dense(self._h, rows, lo, hi, A.dptr(out), max(rows, 1)), self.ctx.h)
        return out[:, :rows].T

    def __getitem__(self, idx):
        """L[j, k] with slices / ints, 0-based (Transfer.jl:183-187)"""
        j, k = idx
        ks = range(k, k + 1) if isinstance(k, int) else range(*k.indices(1 << 30 if k.stop is None else k.stop))
        if len(ks) == 0 or ks.step != 1:
            raise IndexError("column index must be an int or a unit-step slice")
        lo, hi = ks.start, ks.stop
        if isinstance(j, slice) and j.stop is None:
            rows = max(self.colstop(c) for c in range(lo, hi))
        else:
            rows = (j + 1) if isinstance(j, int) else j.stop
        D = self.dense(rows, lo, hi)
        D = D[j] if isinstance(j, int) else D[j, :]
        return D[..., 0] if isinstance(k, int) else D

    def __mul__(self, f):
        """L * Fun (Transfer.jl:58-59 test): applied on the device from the ragged store"""
        if not isinstance(f, Fun):
            return NotImplemented
        c = np.ascontiguousarray(f.coefficients, dtype=np.float64)
        self.resizedata(len(c))
        rows = max([self.colstop(k) for k in range(len(c))] + [1])
        y = np.empty(rows)
        A.check(A.lib().pgb_transfer_apply(self._h, A.dptr(c), len(c), A.dptr(y), rows), self.ctx.h)
        d = self.m.rangedomain
        out = Fun([0.0], (PeriodicSegment if self.rangespace == A.FOURIER else Interval)(d.a, d.b), space=self.rangespace)
        out.coefficients = _chop(y, EPS * max(np.abs(y).max(), 1e-300))
        return out

    def close(self):
        if self._h:
            A.lib().pgb_transfer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _branch_sum_fun(m, g, rangespace=None, ctx=None, maxn=1 << 14):
    """Fun on the range space of y -> sum_b g(v_b(y), w_b(y)), w_b = |v_b'(y)| (product over a composition): the
    inverse branches come from the GPU (K-a, pgb_invert_branches), the closure g is evaluated on the host --
    closures never cross the C ABI -- and the node count doubles until the coefficient tail has decayed, like
    ApproxFun's Fun constructor (Transfer.jl:75-83: Fun(TransferFunction(m, fn), sp))."""
    dm = m if isinstance(m, E.DeviceMap) else E.DeviceMap(m, ctx, None, rangespace)
    rs = dm.desc.range_space
    a, b = dm.desc.ran_a, dm.desc.ran_b
    n = 32
    while True:
        x, v, w = E.invert_branches(dm, n)
        vals = np.zeros(n)
        for c in range(v.shape[0]):
            live = w[c] > 0.0
            if live.any():
                vals[live] += g(v[c][live], w[c][live])
        coef = _transform(rs, vals)
        mx = max(np.abs(coef).max(), 1e-300)
        if np.abs(coef[-max(3, n // 8):]).max() <= 10 * EPS * mx or n >= maxn:
            out = Fun([0.0], (PeriodicSegment if rs == A.FOURIER else Interval)(a, b), space=rs)
            out.coefficients = _chop(coef, 2 * EPS * mx)
            return out
        n *= 2


def transfer(m, fn):
    """transfer(m, fn) = Fun(x -> transferfunction(x, m, fn), rangespace) (Transfer.jl:75-83): L applied to an
    arbitrary host function, sampled through the GPU's inverse branches.  A Fun argument goes through the cached
    matrix instead (L * Fun)."""
    if isinstance(fn, Fun):
        L = m if isinstance(m, Transfer) else Transfer(m)
        return L * fn
    mm = m.m if isinstance(m, Transfer) else m
    return _branch_sum_fun(mm, lambda v, w: w * fn(v))


def lyapunov(obj, r=None):
    """Lyapunov exponent sum(Fun(x -> transferfunction(x, f, x -> log|f'(x)| r(x)), sp)) (statistics.jl:175-198).
    At an inverse-branch point x = v_b(y) the forward derivative is f'(x) = 1/v_b'(y), so the integrand is
    -sum_b w_b log(w_b) r(v_b(y)) with exactly the (v_b, w_b) the inversion kernel returns; for a composition w_b is
    the chain-rule product, which makes this the sum over the stages that statistics.jl:186-198 computes."""
    if isinstance(obj, SolutionInv):
        mm, r = obj.L.m, (obj.acim() if r is None else r)
    else:
        mm = obj.m if isinstance(obj, Transfer) else obj
        r = acim(obj) if r is None else r
    return _branch_sum_fun(mm, lambda v, w: -w * np.log(w) * r(v)).sum()


def _as_transfer(obj):
    if isinstance(obj, Transfer):
        return obj
    if isinstance(obj, AbstractMarkovMap):
        return Transfer(obj)
    raise TypeError(f"cannot make a transfer operator from {type(obj).__name__}")


# ----------------------------------------------------------------------------------------- SolutionInv
class SolutionInv:
    """SolutionInv(L, u=uniform): QR of I - L + (u/sum u) (x) DefiniteIntegral, cached on the device
    (Solution.jl:24-37).  The reference's QR is adaptive on the infinite operator; here the operator is
    truncated to n x n, n doubled until the tail of the solution has decayed (or fixed by `n=`)."""

    def __init__(self, L, u=None, n=None, maxn=8192):
        self.L = _as_transfer(L)
        m = self.L.m
        if not (m.domain == m.rangedomain):
            raise AssertionError("domain(L) == rangedomain(L)")
        self.u = uniform(m.domain) if u is None else u
        self.fixed = n is not None
        self.maxn = maxn
        self.n = 0
        self._h = C.c_void_p()
        self._factor(int(n) if n is not None else max(32, 2 * len(self.u.coefficients)))

    def _factor(self, n):
        if self._h:
            A.lib().pgb_solinv_destroy(self._h)
            self._h = C.c_void_p()
        u = np.ascontiguousarray(self.u.coefficients[:n], dtype=np.float64)
        A.check(A.lib().pgb_solinv_create(self.L._h, n, A.dptr(u), len(u), C.byref(self._h)), self.L.ctx.h)
        self.n = n

    def _solve_n(self, b):
        B = np.zeros(self.n)
        k = min(self.n, len(b))
        B[:k] = b[:k]
        X = np.empty(self.n)
        A.check(A.lib().pgb_solinv_solve(self._h, A.dptr(B), A.dptr(X), 1), self.L.ctx.h)
        return X

    def solve(self, f):
        """K \\ f"""
        b = f.coefficients if isinstance(f, Fun) else np.asarray(f, dtype=np.float64)
        while True:
            x = self._solve_n(b)
            mx = max(np.abs(x).max(), 1e-300)
            tail = np.abs(x[-max(4, self.n // 8):]).max()
            if self.fixed or self.n >= self.maxn or (len(b) <= self.n and tail <= 8 * EPS * mx):
                break
            self._factor(2 * self.n)
        out = self.u._like(x if self.fixed else _chop(x, EPS * mx))
        return out

    def acim(self):
        return self.solve(self.u)

    def close(self):
        if self._h:
            A.lib().pgb_solinv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_solinv(obj):
    return obj if isinstance(obj, SolutionInv) else SolutionInv(obj)


def acim(obj, u=None):
    """acim(K) = K \\ u  (statistics.jl:9-11)"""
    if isinstance(obj, SolutionInv):
        return obj.acim()
    return SolutionInv(obj, u).acim()


def _as_fun(K, X):
    if isinstance(X, Fun):
        return X
    return Fun(X, K.L.m.domain, space=K.L.domainspace)


def linearresponse(obj, X):
    """K \\ (-(acim(K) * X)')  (statistics.jl:17-20)"""
    K = _as_solinv(obj)
    X = _as_fun(K, X)
    if K.L.domainspace == A.CHEBYSHEV:
        ends = X(np.array([X.a, X.b]))
        if not np.all(np.abs(ends) <= np.sqrt(EPS * (X.b - X.a))):
            raise AssertionError("X must vanish at the boundary of its domain (statistics.jl:18)")
    return K.solve(-((K.acim() * X).derivative()))


def zero_to(Af, rho=None):
    """rho * (A - int A d rho)  (statistics.jl:27-30)"""
    rho = uniform(Af) if rho is None else rho
    rA = rho * Af
    return rA - rho * (rA.sum() / rho.sum())


def correlationsum(obj, Af, rho=None):
    """K \\ zero_to(A, rho)  (statistics.jl:36)"""
    K = _as_solinv(obj)
    Af = _as_fun(K, Af)
    return K.solve(zero_to(Af, K.acim() if rho is None else rho))


def birkhoffcov(obj, Af, Bf):
    """statistics.jl:42-45"""
    K = _as_solinv(obj)
    Af, Bf = _as_fun(K, Af), _as_fun(K, Bf)
    rho = K.acim()
    Az, Bz = zero_to(Af, rho), zero_to(Bf, rho)
    return (Bf * K.solve(Az)).sum() + (Af * K.solve(Bz)).sum() - (Bf * Az).sum()


def birkhoffvar(obj, Af):
    """2 sum(A (K \\ Az)) - sum(A Az)  (statistics.jl:51-54)"""
    K = _as_solinv(obj)
    Af = _as_fun(K, Af)
    Az = zero_to(Af, K.acim())
    return 2 * (Af * K.solve(Az)).sum() - (Af * Az).sum()


def _chopfun(f):
    """chop!(f): drop the trailing coefficients at rounding level (ApproxFun's default tolerance)"""
    c = f.coefficients
    return f._like(_chop(c, 10 * EPS * max(np.abs(c).max(), 1e-300)))


def covariancefunction(obj, Af, Bf=None, n=None, r=None, tol=None):
    """lag-covariance function by iterating the cached operator (statistics.jl:66-168): entry k is the expectation of
    A . A o f^k (one observable) or the pair (E[B . A o f^k], E[A . B o f^k]) (two).  n fixed, or doubled from 16
    until the last third of the sequence is below tol (default eps * |A|^2, resp. eps * |A| |B|).  Every L * Fun is a
    device mat-vec on the ragged store (pgb_transfer_apply)."""
    K = obj if isinstance(obj, SolutionInv) else None
    L = K.L if K is not None else _as_transfer(obj)
    if r is None:
        r = (K or SolutionInv(L)).acim()
    dom = L.m.domain
    Af = Af if isinstance(Af, Fun) else Fun(Af, dom, space=L.domainspace)
    if Bf is not None and not isinstance(Bf, Fun):
        Bf = Fun(Bf, dom, space=L.domainspace)
    rsum = r.sum()

    def advance(z):
        z = L * z
        z = z - r * (z.sum() / rsum)
        return _chopfun(z)

    if Bf is None:
        Az = zero_to(Af, r)
        cf = [(Af * Az).sum()]

        def extend(upto):
            nonlocal Az
            while len(cf) < upto + 1:
                Az = advance(Az)
                cf.append((Af * Az).sum())
        if n is not None:
            extend(n)
            return np.array(cf)
        tol = EPS * float(np.linalg.norm(Af.coefficients)) ** 2 if tol is None else tol
        nn = 16
        extend(nn)
        while np.abs(np.array(cf[(2 * nn) // 3 - 1:])).max() > tol and nn <= (1 << 20):
            nn *= 2
            extend(nn)
        return _chop(np.array(cf), EPS)
    Az, Bz = zero_to(Af, r), zero_to(Bf, r)
    cfA, cfB = [(Af * Bz).sum()], [(Af * Bz).sum()]

    def extend2(upto):
        nonlocal Az, Bz
        while len(cfA) < upto + 1:
            Az, Bz = advance(Az), advance(Bz)
            cfA.append((Af * Bz).sum())
            cfB.append((Bf * Az).sum())
    if n is not None:
        extend2(n)
        return np.array(cfA), np.array(cfB)
    tol = EPS * float(np.linalg.norm(Af.coefficients) * np.linalg.norm(Bf.coefficients)) if tol is None else tol
    nn = 16
    extend2(nn)
    while max(np.abs(np.array(cfA[(2 * nn) // 3 - 1:])).max(), np.abs(np.array(cfB[(2 * nn) // 3 - 1:])).max()) > tol and nn <= (1 << 20):
        nn *= 2
        extend2(nn)
    return _chop(np.array(cfA), EPS), _chop(np.array(cfB), EPS)


def eigvals(obj, n):
    """eigvals(L, n): eigenvalues of the leading n x n block, sorted by decreasing modulus (poetry.jl:56)"""
    L = _as_transfer(obj)
    wr, wi = np.empty(n), np.empty(n)
    A.check(A.lib().pgb_eigvals(L._h, int(n), A.dptr(wr), A.dptr(wi)), L.ctx.h)
    ev = wr + 1j * wi
    return ev[np.argsort(-np.abs(ev), kind="stable")]


# ----------------------------------------------------------------------------------------- batched sweeps
class BatchSweep:
    """Many maps of one shape (e.g. perturb(M, X, eps_i), poetry.jl:46): one fixed-grid n x n fill and one
    acim solve per map, all in a handful of launches (pgb_batch_*)."""

    def __init__(self, maps, ctx=None):
        self.ctx = ctx or E.get_context()
        self.descs = [m.to_desc() for m in maps]
        self.nmaps = len(maps)
        arr = (A.pgb_map_desc * self.nmaps)(*self.descs)
        self._arr = arr
        self._h = C.c_void_p()
        A.check(A.lib().pgb_batch_create(self.ctx.h, arr, self.nmaps, C.byref(self._h)), self.ctx.h)
        d = self.descs[0]
        self.space, self.a, self.b = d.domain_space, d.dom_a, d.dom_b

    def fill_acim(self, n_nodes, n, want_L=False):
        """-> (rho[nmaps, n], L[nmaps, n(rows), n(cols)] or None)"""
        rho = np.empty((self.nmaps, n))
        Lh = np.empty((self.nmaps, n, n)) if want_L else None
        A.check(A.lib().pgb_batch_fill_acim(self._h, n_nodes, n, A.dptr(Lh), A.dptr(rho)), self.ctx.h)
        return rho, (None if Lh is None else Lh.transpose(0, 2, 1))

    def eigvals(self, n_nodes, n, nev=8):
        """eigvals(L_i, n) for every map of the sweep (poetry.jl:56-66): the nev eigenvalues of largest modulus, descending
        -> complex array [nmaps, nev]"""
        wr, wi = np.empty((self.nmaps, nev)), np.empty((self.nmaps, nev))
        A.check(A.lib().pgb_batch_eigvals(self._h, n_nodes, n, nev, A.dptr(wr), A.dptr(wi)), self.ctx.h)
        return wr + 1j * wi

    def close(self):
        if self._h:
            A.lib().pgb_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
