"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/_build/liboracle.so plus the dense
(numpy/LAPACK) restatement of the solves that consume the matrix.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (poltergeist.jl_b200/) never does.

Restated reference code (paths relative to the reference checkout):
  getindex / dense / invert      -> oracle.c (src/spectral/Transfer.jl:89-181 etc.)
  solinv_matrix / acim           -> src/spectral/Solution.jl:17-37, src/spectral/statistics.jl:9-11
  linearresponse                 -> src/spectral/statistics.jl:17-20
  zero_to / birkhoffvar          -> src/spectral/statistics.jl:27-30,51-54
  lyapunov                       -> src/spectral/statistics.jl:175-177
  eigvals                        -> src/poetry.jl:56
Parity status: conventions pinned by the reference's known-answer tests
(tests/test_oracle_pins.py); individual matrix entries are otherwise "parity unpinned"
(the reference ships no stored matrices and cannot run here: no Julia).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
CHEBYSHEV, FOURIER = 0, 1
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_generic.inc")]
    src.append(os.path.join(_HERE, "..", "include", "poltergeist_b200.h"))
    if force or not os.path.exists(_SO) or (
            all(os.path.exists(s) for s in src) and any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, dp, ip = C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int32)
        i64p = C.POINTER(C.c_int64)
        L.orc_ref_getindex.argtypes = [vp, ip, i64, i64, i64, dp, i64, i64p, i64p, ip, i64p]
        L.orc_ref_dense.argtypes = [vp, i64, i64, i64, i64, dp, i64]
        L.orc_ref_invert.argtypes = [vp, i64, dp, dp, dp, i64p]
        L.orc_ref_points.argtypes = [C.c_int, C.c_double, C.c_double, i64, dp]
        L.orc_ref_transform.argtypes = [C.c_int, dp, i64, dp]
        L.orc_ref_fun_eval.argtypes = [C.c_int, C.c_double, C.c_double, dp, i64, C.c_double]
        L.orc_ref_fun_eval.restype = C.c_double
        L.orc_ref_transferfunction.argtypes = [vp, C.c_double, i64, C.POINTER(C.c_int)]
        L.orc_ref_transferfunction.restype = C.c_double
        L.orc_exact_invert.argtypes = [vp, i64, dp, dp, dp, C.c_int]
        L.orc_exact_dense.argtypes = [vp, i64, i64, i64, i64, dp, i64, C.c_int]
        L.orc_exact_transfer_fun.argtypes = [vp, dp, i64, dp, i64, dp, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _desc_ptr(desc):
    return C.cast(C.byref(desc), C.c_void_p)


def _ncomp(desc):
    if desc.npath > 0:
        return desc.npath
    n = 1
    for s in range(desc.nstage):
        n *= desc.stages[s].nbranch
    return n


class OracleError(RuntimeError):
    pass


def _chk(rc):
    if rc == 3:
        raise OracleError("Newton: failure to converge")
    if rc:
        raise OracleError(f"oracle status {rc}")


def set_checkpoints(cheb=None, fourier=None):
    """move ApproxFun.checkpoints (None = the recalled defaults): robustness tests only"""
    L = lib()
    c = None if cheb is None else (C.c_double * 3)(*cheb)
    f = None if fourier is None else (C.c_double * 3)(*fourier)
    L.orc_set_checkpoints(c, f)


def points(space, a, b, n):
    x = np.empty(n)
    lib().orc_ref_points(space, a, b, n, _dp(x))
    return x


def transform(space, vals):
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    c = np.empty_like(vals)
    lib().orc_ref_transform(space, _dp(vals), len(vals), _dp(c))
    return c


def fun_eval(space, a, b, coeffs, x):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
    return np.array([lib().orc_ref_fun_eval(space, a, b, _dp(coeffs), len(coeffs), float(xx))
                     for xx in np.atleast_1d(x)])


class RefTransfer:
    """The faithful cached operator: ConcreteTransfer.colstops memo + RaggedMatrix cache,
    grown by `resizedata` exactly like Transfer.jl:201-217 (one transfer_getindex per call)."""

    def __init__(self, desc, max_cols=1 << 16):
        self.desc = desc
        self.colstops = np.full(max_cols, -1, dtype=np.int32)
        self.data = np.empty(0)
        self.cols = np.array([1], dtype=np.int64)
        self.m = 0
        self.ncols = 0
        self.nodes_used = np.empty(0, dtype=np.int32)
        self.newton_iters = 0

    def resizedata(self, n):
        if n <= self.ncols:
            return self
        lo, hi = self.ncols, n
        cap = 1 << 16
        while True:
            data = np.empty(cap)
            cols = np.empty(hi - lo + 1, dtype=np.int64)
            nodes = np.zeros(hi - lo, dtype=np.int32)
            m = C.c_int64(0)
            it = C.c_int64(0)
            saved = self.colstops.copy()
            rc = lib().orc_ref_getindex(_desc_ptr(self.desc), self.colstops.ctypes.data_as(C.POINTER(C.c_int32)),
                                        lo, hi, 0, _dp(data), cap, cols.ctypes.data_as(C.POINTER(C.c_int64)),
                                        C.byref(m), nodes.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(it))
            if rc == 1:
                self.colstops = saved
                cap *= 4
                continue
            _chk(rc)
            break
        nnz = cols[-1] - 1
        self.data = np.concatenate([self.data, data[:nnz]])
        self.cols = np.concatenate([self.cols, cols[1:] + (self.cols[-1] - 1)])
        self.m = max(self.m, m.value)
        self.ncols = n
        self.nodes_used = np.concatenate([self.nodes_used, nodes])
        self.newton_iters += it.value
        return self

    def colstop(self, k):
        """0-based column"""
        self.resizedata(k + 1)
        return int(self.cols[k + 1] - self.cols[k])

    def dense(self, rows, ncols):
        self.resizedata(ncols)
        out = np.zeros((rows, ncols))
        for j in range(ncols):
            c = self.data[self.cols[j] - 1:self.cols[j + 1] - 1]
            r = min(rows, len(c))
            out[:r, j] = c[:r]
        return out


def ref_dense(desc, n, col_lo, col_hi, rows):
    out = np.zeros((col_hi - col_lo, rows))
    _chk(lib().orc_ref_dense(_desc_ptr(desc), n, col_lo, col_hi, rows, _dp(out), rows))
    return out.T  # (rows, cols) view of the column-major block


def exact_dense(desc, n, col_lo, col_hi, rows, prec=0):
    out = np.zeros((col_hi - col_lo, rows))
    _chk(lib().orc_exact_dense(_desc_ptr(desc), n, col_lo, col_hi, rows, _dp(out), rows, prec))
    return out.T


def ref_invert(desc, n):
    nc = _ncomp(desc)
    x, v, w = np.empty(n), np.empty((nc, n)), np.empty((nc, n))
    it = C.c_int64(0)
    _chk(lib().orc_ref_invert(_desc_ptr(desc), n, _dp(x), _dp(v), _dp(w), C.byref(it)))
    return x, v, w, it.value


def exact_invert(desc, n, prec=0):
    nc = _ncomp(desc)
    x, v, w = np.empty(n), np.empty((nc, n)), np.empty((nc, n))
    _chk(lib().orc_exact_invert(_desc_ptr(desc), n, _dp(x), _dp(v), _dp(w), prec))
    return x, v, w


def exact_transfer_fun(desc, coeffs, xs, prec=0):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    out = np.empty_like(xs)
    _chk(lib().orc_exact_transfer_fun(_desc_ptr(desc), _dp(coeffs), len(coeffs), _dp(xs), len(xs), _dp(out), prec))
    return out


def ref_transferfunction(desc, x, col):
    err = C.c_int(0)
    y = lib().orc_ref_transferfunction(_desc_ptr(desc), float(x), col, C.byref(err))
    if err.value:
        raise OracleError("Newton: failure to converge")
    return y


# ---------------------------------------------------------------------------------
# dense restatement of Solution.jl / statistics.jl on an n x n truncation
# ---------------------------------------------------------------------------------
def integral_functional(space, a, b, n):
    """DefiniteIntegral(space)[1:n]."""
    w = np.zeros(n)
    if space == CHEBYSHEV:
        k = np.arange(0, n, 2)
        w[0::2] = (b - a) / 2 * 2.0 / (1.0 - k.astype(float) ** 2)
    else:
        w[0] = b - a
    return w


def uniform(space, a, b, n):
    """uniform(S): the constant 1/|d| (Solution.jl:17-22)."""
    u = np.zeros(n)
    u[0] = 1.0 / (b - a)
    return u


def solinv_matrix(L, space, a, b, u=None):
    """I - L + (u/sum(u)) (x) DefiniteIntegral  (Solution.jl:36)."""
    n = L.shape[0]
    w = integral_functional(space, a, b, n)
    u = uniform(space, a, b, n) if u is None else np.pad(u, (0, n - len(u)))
    return np.eye(n) - L + np.outer(u / (w @ u), w)


class RefSolutionInv:
    def __init__(self, L, space, a, b, u=None):
        import scipy.linalg as sla
        self.space, self.a, self.b, self.n = space, a, b, L.shape[0]
        self.u = uniform(space, a, b, self.n) if u is None else np.pad(u, (0, self.n - len(u)))
        self.Q, self.R = sla.qr(solinv_matrix(L, space, a, b, u))

    def solve(self, rhs):
        import scipy.linalg as sla
        rhs = np.pad(rhs, (0, self.n - len(rhs)))
        return sla.solve_triangular(self.R, self.Q.T @ rhs)

    def acim(self):
        return self.solve(self.u)


# --- minimal Fun algebra, independent of the product's (numpy.polynomial / FFT based) ----
def cheb_to_std(a, b, c):
    return np.asarray(c, dtype=float)


def fun_mul(space, f, g, n):
    """coefficients of f*g truncated/padded to n"""
    if space == CHEBYSHEV:
        p = np.polynomial.chebyshev.chebmul(f, g)
    else:
        m = 1
        while m < 2 * (len(f) + len(g)) + 2:
            m *= 2
        th = 2 * np.pi * np.arange(m) / m
        p = transform(FOURIER, _fourier_eval_theta(f, th) * _fourier_eval_theta(g, th))
    out = np.zeros(n)
    out[:min(n, len(p))] = p[:n]
    return out


def _fourier_eval_theta(c, th):
    s = np.full_like(th, c[0] if len(c) else 0.0)
    for k in range(1, len(c)):
        mm = (k + 1) // 2
        s += c[k] * (np.sin(mm * th) if k & 1 else np.cos(mm * th))
    return s


def fun_diff(space, a, b, f):
    if space == CHEBYSHEV:
        return np.polynomial.chebyshev.chebder(f) * (2.0 / (b - a)) if len(f) > 1 else np.zeros(1)
    d = np.zeros(len(f) + (len(f) % 2 == 0))
    sc = 2 * np.pi / (b - a)
    for k in range(1, len(f)):
        mm = (k + 1) // 2
        if k & 1:    # sin(m th) -> m cos(m th): slot 2m
            d[2 * mm] += sc * mm * f[k]
        else:        # cos(m th) -> -m sin(m th): slot 2m-1
            d[2 * mm - 1] -= sc * mm * f[k]
    return d


def fun_sum(space, a, b, f):
    return float(integral_functional(space, a, b, len(f)) @ f)


def linearresponse(K, X):
    """K \\ (-(acim(K) X)')  (statistics.jl:17-20); X as coefficients"""
    rho = K.acim()
    rx = fun_mul(K.space, rho, X, K.n + len(X))
    return K.solve(-fun_diff(K.space, K.a, K.b, rx)[:K.n])


def zero_to(space, a, b, Af, rho, n):
    """rho*(A - int A drho) (statistics.jl:27-30)"""
    rA = fun_mul(space, rho, Af, n)
    return rA - np.pad(rho, (0, n - len(rho)))[:n] * fun_sum(space, a, b, rA) / fun_sum(space, a, b, rho)


def birkhoffvar(K, Af):
    """2 sum(A (K\\Az)) - sum(A Az) (statistics.jl:51-54)"""
    rho = K.acim()
    Az = zero_to(K.space, K.a, K.b, Af, rho, K.n)
    s = K.solve(Az)
    m = K.n + len(Af)
    return 2 * fun_sum(K.space, K.a, K.b, fun_mul(K.space, Af, s, m)) - fun_sum(K.space, K.a, K.b, fun_mul(K.space, Af, Az, m))


def eigvals(L, n):
    """eigvals(L, n): dense eigenvalues of the leading n x n block (poetry.jl:56)."""
    ev = np.linalg.eigvals(L[:n, :n])
    return ev[np.argsort(-np.abs(ev))]
