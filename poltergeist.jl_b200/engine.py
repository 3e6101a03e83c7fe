"""Thin object layer over the C ABI (include/poltergeist_b200.h): context, device-resident maps,
the fixed-grid fill and the dense solves.  Everything here calls libpoltergeist_b200.so through
ctypes; nothing computes on the CPU and nothing imports the test oracle.

Reference interfaces mirrored (paths relative to the reference checkout):
  invert_branches   mapinvP over all branches at points(rangespace, n)
                    (src/maps/ExpandingBranch.jl:34-37,66; src/maps/CircleMap.jl:31-37,78-83)
  transfer_fill     transfer_getindex on a fixed grid (src/spectral/Transfer.jl:95-181)
  DenseSolutionInv  SolutionInv / `\\` / acim (src/spectral/Solution.jl:24-37, statistics.jl:9)
  eigvals_dense     eigvals(L, n) (src/poetry.jl:56)
"""
import ctypes as C

import numpy as np

from . import _abi as A

_contexts = {}


class Context:
    """pgb_ctx: one per (process, device)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._lib = A.lib()
        A.check(self._lib.pgb_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    @property
    def h(self):
        return self._h

    def sync(self):
        A.check(self._lib.pgb_ctx_sync(self._h), self._h)

    def launch_count(self):
        return int(self._lib.pgb_ctx_launch_count(self._h))

    def set_stream(self, cuda_stream_ptr):
        A.check(self._lib.pgb_ctx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), self._h)

    def timer_start(self):
        A.check(self._lib.pgb_ctx_timer_start(self._h), self._h)

    def timer_stop(self):
        ms = C.c_double(0.0)
        A.check(self._lib.pgb_ctx_timer_stop(self._h, C.byref(ms)), self._h)
        return ms.value

    def fp64_peak_tflops(self):
        tf = C.c_double(0.0)
        A.check(self._lib.pgb_bench_fp64_peak(self._h, C.byref(tf)), self._h)
        return tf.value

    def profile(self, enable):
        A.check(self._lib.pgb_ctx_profile(self._h, 1 if enable else 0), self._h)

    def profile_read(self, kernel_class):
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        A.check(self._lib.pgb_ctx_profile_read(self._h, int(kernel_class), C.byref(ms), C.byref(cnt)), self._h)
        return ms.value, cnt.value

    def graph_capture(self, fn):
        """run fn() with the context stream in capture mode -> replayable graph handle (everything fn enqueues must
        be stream-ordered and allocation-free: warm the path up first)"""
        A.check(self._lib.pgb_ctx_graph_begin(self._h), self._h)
        try:
            fn()
        finally:
            g = C.c_void_p()
            rc = self._lib.pgb_ctx_graph_end(self._h, C.byref(g))
        A.check(rc, self._h)
        return g

    def graph_launch(self, g):
        A.check(self._lib.pgb_graph_launch(self._h, g), self._h)

    def close(self):
        if self._h:
            self._lib.pgb_ctx_destroy(self._h)
            self._h = C.c_void_p()


def get_context(device=0):
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


K_INVERT, K_BASIS, K_TRANSFORM = 0, 1, 2


class DeviceBuffer:
    """cudaMalloc'ed block owned by this object."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self._p = C.c_void_p()
        A.check(A.lib().pgb_dev_alloc(ctx.h, self.nbytes, C.byref(self._p)), ctx.h)

    @property
    def ptr(self):
        return self._p

    def download(self, dtype=np.float64, count=None, offset_bytes=0):
        count = (self.nbytes - offset_bytes) // np.dtype(dtype).itemsize if count is None else count
        out = np.empty(count, dtype=dtype)
        A.check(A.lib().pgb_dev_download(self.ctx.h, out.ctypes.data_as(C.c_void_p),
                                         C.c_void_p(self._p.value + offset_bytes), out.nbytes), self.ctx.h)
        return out

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        A.check(A.lib().pgb_dev_upload(self.ctx.h, self._p, arr.ctypes.data_as(C.c_void_p), arr.nbytes), self.ctx.h)

    def free(self):
        if self._p:
            A.lib().pgb_dev_free(self.ctx.h, self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceMap:
    """pgb_map: the device-resident compiled descriptor of a map (+ spaces)."""

    def __init__(self, m, ctx=None, domain_space=None, range_space=None):
        self.ctx = ctx or get_context()
        self.desc = m.to_desc(domain_space, range_space) if hasattr(m, "to_desc") else m
        self._h = C.c_void_p()
        A.check(A.lib().pgb_map_create(self.ctx.h, C.byref(self.desc), C.byref(self._h)), self.ctx.h)

    @property
    def h(self):
        return self._h

    @property
    def ncomposite(self):
        return int(A.lib().pgb_map_ncomposite(self._h))

    def invalidate(self):
        """forget the memoised inverse branches (benchmarks: time the whole path every step)"""
        A.check(A.lib().pgb_map_invalidate(self._h), self.ctx.h)

    def close(self):
        if self._h:
            A.lib().pgb_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_devmap(m, ctx=None):
    return m if isinstance(m, DeviceMap) else DeviceMap(m, ctx)


def _opts(chop, tol):
    o = A.pgb_chop_opts()
    o.tol = 0.0 if tol is None else float(tol)
    o.chop = 1 if chop else 0
    return o


def invert_branches(m, n):
    """(x, v, w): nodes, inverse-branch values and prod |v'| at the n collocation nodes, [ncomp, n]."""
    dm = _as_devmap(m)
    nc = dm.ncomposite
    x, v, w = np.empty(n), np.empty((nc, n)), np.empty((nc, n))
    A.check(A.lib().pgb_invert_branches(dm.ctx.h, dm.h, n, A.dptr(x), A.dptr(v), A.dptr(w)), dm.ctx.h)
    return x, v, w


def transfer_fill(m, n, col_lo, col_hi, rows=None, chop=False, tol=None):
    """Columns [col_lo, col_hi) of the transfer matrix on the fixed n-node grid, host result.
    Returns (block[rows, ncols], colstops[ncols])."""
    dm = _as_devmap(m)
    rows = n if rows is None else rows
    ncols = col_hi - col_lo
    out = np.zeros((ncols, rows))  # column-major block == C-order (ncols, rows)
    cs = np.zeros(ncols, dtype=np.int32)
    o = _opts(chop, tol)
    A.check(A.lib().pgb_transfer_fill(dm.ctx.h, dm.h, n, col_lo, col_hi, rows, A.dptr(out), rows,
                                      cs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(o)), dm.ctx.h)
    return out.T, cs


def transfer_fill_ragged(m, n, col_lo, col_hi, tol=None, tail=False):
    """The same fixed-grid fill returned as the reference's RaggedMatrix(data, cols, m): chop!, interior zeroing
    and colstops on the device, only retained entries copied to the host.  -> (data, cols[1-based], colstops, m).
    tail=True: pgb_transfer_fill_ragged_tail -- the library packs against the end of its buffer and produces the
    columns in descending order (the long columns cross PCIe first); `data` is the view of the retained entries."""
    dm = _as_devmap(m)
    ncols = col_hi - col_lo
    cols = np.zeros(ncols + 1, dtype=np.int64)
    cs = np.zeros(max(ncols, 1), dtype=np.int32)
    o = _opts(True, tol)
    mm, nnz, off = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    cap = max(1, min(ncols * n, 1 << 24))
    colsp, csp = cols.ctypes.data_as(C.POINTER(C.c_int64)), cs.ctypes.data_as(C.POINTER(C.c_int32))
    while True:
        data = np.empty(cap)
        if tail:
            rc = A.lib().pgb_transfer_fill_ragged_tail(dm.ctx.h, dm.h, n, col_lo, col_hi, A.dptr(data), cap, C.byref(off), colsp, csp,
                                                       C.byref(mm), C.byref(nnz), C.byref(o))
        else:
            rc = A.lib().pgb_transfer_fill_ragged(dm.ctx.h, dm.h, n, col_lo, col_hi, A.dptr(data), cap, colsp, csp,
                                                  C.byref(mm), C.byref(nnz), C.byref(o))
        if rc == A.PGB_ERR_BAD_ARG and nnz.value > cap:
            cap = nnz.value
            continue
        A.check(rc, dm.ctx.h)
        return data[off.value:off.value + nnz.value], cols, cs[:ncols], mm.value


class DeviceBlock:
    """A dense column-major block of the transfer matrix resident in HBM."""

    def __init__(self, ctx, rows, ncols):
        self.ctx, self.rows, self.ncols, self.ld = ctx, rows, ncols, rows
        self.buf = DeviceBuffer(ctx, 8 * rows * ncols)
        self.colstops = DeviceBuffer(ctx, 4 * max(ncols, 1))

    def to_host(self):
        return self.buf.download(np.float64).reshape(self.ncols, self.rows).T

    def colstops_host(self):
        return self.colstops.download(np.int32, self.ncols)


def transfer_fill_device(m, n, col_lo, col_hi, rows=None, chop=False, tol=None, out=None, col_offset=0, sparse=False):
    """Same fill, block left in HBM (asynchronous on the context stream).  `out`: an existing
    DeviceBlock to write into at column offset `col_offset` (sharded fills).
    sparse (chop only): pgb_transfer_fill_device_sparse with out.colstops as the record of what the block holds --
    only rows below max(new, old colstop) are written.  A fresh block starts with colstops = rows (nothing known)."""
    dm = _as_devmap(m)
    rows = n if rows is None else rows
    ncols = col_hi - col_lo
    if out is None:
        out = DeviceBlock(dm.ctx, rows, ncols)
        col_offset = 0
        if sparse:
            out.colstops.upload(np.full(ncols, rows, dtype=np.int32))
    o = _opts(chop, tol)
    dst = C.c_void_p(out.buf.ptr.value + 8 * out.ld * col_offset)
    cdst = C.c_void_p(out.colstops.ptr.value + 4 * col_offset)
    if sparse:
        A.check(A.lib().pgb_transfer_fill_device_sparse(dm.ctx.h, dm.h, n, col_lo, col_hi, rows, dst, out.ld, cdst, cdst, C.byref(o)),
                dm.ctx.h)
    else:
        A.check(A.lib().pgb_transfer_fill_device(dm.ctx.h, dm.h, n, col_lo, col_hi, rows, dst, out.ld, cdst, C.byref(o)),
                dm.ctx.h)
    return out


class DenseSolutionInv:
    """QR of the n x n truncation of I - L + (u/sum u) (x) DefiniteIntegral, kept on the device."""

    def __init__(self, block, n, space, dom_a, dom_b, u=None, structured=False):
        """structured: the block came from a chopped fill (chop=True) whose colstops are in block.colstops --
        the factorisation then skips the part of the matrix that is already upper triangular."""
        self.ctx, self.n, self.space, self.a, self.b = block.ctx, n, space, dom_a, dom_b
        self._h = C.c_void_p()
        up = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        cs = block.colstops.ptr if structured else C.c_void_p()
        A.check(A.lib().pgb_solinv_create_dense(self.ctx.h, block.buf.ptr, block.ld, n, cs, space, dom_a, dom_b,
                                                A.dptr(up), 0 if up is None else len(up), C.byref(self._h)), self.ctx.h)

    @property
    def factored_block(self):
        return int(A.lib().pgb_solinv_factored_block(self._h))

    def solve(self, rhs):
        rhs = np.asarray(rhs, dtype=np.float64)
        one = rhs.ndim == 1
        B = np.zeros((1 if one else rhs.shape[1], self.n))
        if one:
            B[0, :min(self.n, len(rhs))] = rhs[:self.n]
        else:
            B[:, :min(self.n, rhs.shape[0])] = rhs[:self.n].T
        X = np.empty_like(B)
        A.check(A.lib().pgb_solinv_solve(self._h, A.dptr(B), A.dptr(X), B.shape[0]), self.ctx.h)
        return X[0] if one else X.T

    def acim(self):
        rho = np.empty(self.n)
        A.check(A.lib().pgb_solinv_acim(self._h, A.dptr(rho)), self.ctx.h)
        return rho

    def close(self):
        if self._h:
            A.lib().pgb_solinv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def eigvals_dense(block, n):
    """eigenvalues of the leading n x n block, sorted by decreasing modulus."""
    wr, wi = np.empty(n), np.empty(n)
    A.check(A.lib().pgb_eigvals_dense(block.ctx.h, block.buf.ptr, block.ld, n, A.dptr(wr), A.dptr(wi)), block.ctx.h)
    ev = wr + 1j * wi
    return ev[np.argsort(-np.abs(ev), kind="stable")]


# ----------------------------------------------------------------------------------------------- multi-GPU, one process
class _BorrowedContext:
    """view of a context owned by a pgb_group (never destroyed from here)"""

    def __init__(self, h, device):
        self._h, self.device = C.c_void_p(h), device

    @property
    def h(self):
        return self._h

    def sync(self):
        A.check(A.lib().pgb_ctx_sync(self._h), self._h)

    def launch_count(self):
        return int(A.lib().pgb_ctx_launch_count(self._h))


class _Ptr:
    def __init__(self, p):
        self.ptr = C.c_void_p(p)


class GroupMatrix:
    """pgb_gmatrix: a full copy of the rows x ncols matrix (+ colstops) on every device of the group, in symmetric
    memory allocated by the library (peer-addressable; NVSwitch multicast where available)."""

    def __init__(self, group, rows, ncols, multicast=True):
        self.group, self.rows, self.ncols, self.ld = group, rows, ncols, rows
        self._h = C.c_void_p()
        A.check(A.lib().pgb_gmatrix_create(group.h, rows, ncols, 1 if multicast else 0, C.byref(self._h)), group.ctx(0).h)

    @property
    def h(self):
        return self._h

    @property
    def multicast(self):
        return bool(A.lib().pgb_gmatrix_multicast(self._h))

    def block(self, rank):
        """DeviceBlock-like view of copy `rank` (for DenseSolutionInv / eigvals_dense on that device)"""
        b = _Ptr(0)
        b.ctx, b.ld, b.rows, b.ncols = self.group.ctx(rank), self.ld, self.rows, self.ncols
        b.buf = _Ptr(A.lib().pgb_gmatrix_ptr(self._h, rank))
        b.colstops = _Ptr(A.lib().pgb_gmatrix_colstops(self._h, rank))
        return b

    def to_host(self, rank):
        out = np.empty((self.ncols, self.rows))
        ctx = self.group.ctx(rank)
        A.check(A.lib().pgb_dev_download(ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(A.lib().pgb_gmatrix_ptr(self._h, rank)), out.nbytes), ctx.h)
        return out.T

    def colstops_host(self, rank):
        out = np.empty(self.ncols, dtype=np.int32)
        ctx = self.group.ctx(rank)
        A.check(A.lib().pgb_dev_download(ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(A.lib().pgb_gmatrix_colstops(self._h, rank)), out.nbytes), ctx.h)
        return out

    def close(self):
        if self._h:
            A.lib().pgb_gmatrix_destroy(self._h)
            self._h = C.c_void_p()


class Group:
    """pgb_group: several GPUs driven from this one process (the reference is a single Julia task)."""

    def __init__(self, devices):
        self.devices = [int(d) for d in devices]
        self._h = C.c_void_p()
        arr = (C.c_int * len(self.devices))(*self.devices)
        A.check(A.lib().pgb_group_create(arr, len(self.devices), C.byref(self._h)))
        self._ctx = [_BorrowedContext(A.lib().pgb_group_ctx(self._h, r), d) for r, d in enumerate(self.devices)]
        self._pinned = []

    @property
    def h(self):
        return self._h

    @property
    def size(self):
        return len(self.devices)

    def ctx(self, rank):
        return self._ctx[rank]

    def sync(self):
        A.check(A.lib().pgb_group_sync(self._h), self._ctx[0].h)

    def map(self, m):
        desc = m.to_desc() if hasattr(m, "to_desc") else m
        h = C.c_void_p()
        A.check(A.lib().pgb_group_map_create(self._h, C.byref(desc), C.byref(h)), self._ctx[0].h)
        gm = _Ptr(h.value)
        gm.h, gm.desc = h, desc
        return gm

    def invalidate(self, gm):
        for r in range(self.size):
            A.check(A.lib().pgb_map_invalidate(C.c_void_p(A.lib().pgb_group_map(gm.h, r))), self._ctx[r].h)

    def matrix(self, rows, ncols, multicast=True):
        return GroupMatrix(self, rows, ncols, multicast)

    def fill(self, gm, n, M, chop=True, tol=None):
        o = _opts(chop, tol)
        A.check(A.lib().pgb_group_fill(self._h, gm.h, n, M.h, C.byref(o)), self._ctx[0].h)

    def shard(self, ncols, rank):
        lo, hi = C.c_int64(0), C.c_int64(0)
        A.check(A.lib().pgb_column_shard(ncols, self.size, rank, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def pinned(self, count, dtype=np.float64):
        """page-locked numpy array (pgb_host_alloc), kept alive by the group"""
        nbytes = int(count) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        A.check(A.lib().pgb_host_alloc(nbytes, C.byref(p)))
        self._pinned.append(p)
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(count))

    def fill_ragged(self, gm, n, col_lo, col_hi, tol=None, data=None):
        """one RaggedMatrix(data, cols, m) in host memory from all devices -> (data, cols[1-based], colstops, m)"""
        ncols = col_hi - col_lo
        cols = np.zeros(ncols + 1, dtype=np.int64)
        cs = np.zeros(max(ncols, 1), dtype=np.int32)
        o = _opts(True, tol)
        mm, nnz = C.c_int64(0), C.c_int64(0)
        if data is None:
            data = self.pinned(max(1, min(ncols * n, 1 << 24)))
        while True:
            rc = A.lib().pgb_group_fill_ragged(self._h, gm.h, n, col_lo, col_hi, A.dptr(data), len(data), cols.ctypes.data_as(C.POINTER(C.c_int64)),
                                               cs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(mm), C.byref(nnz), C.byref(o))
            if rc == A.PGB_ERR_BAD_ARG and nnz.value > len(data):
                data = self.pinned(nnz.value)
                continue
            A.check(rc, self._ctx[0].h)
            return data[:nnz.value], cols, cs[:ncols], mm.value

    def close(self):
        for p in self._pinned:
            A.lib().pgb_host_free(p)
        self._pinned = []
        if self._h:
            A.lib().pgb_group_destroy(self._h)
            self._h = C.c_void_p()
