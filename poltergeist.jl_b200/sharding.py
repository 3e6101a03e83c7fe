"""Column sharding of the transfer-matrix fill across one-process-per-GPU ranks.

Mode columns are independent given the inverse branches, so rank r of W fills a contiguous block of columns
straight into its slice of the full column-major N x rows buffer and the shards are reassembled on every rank
with one all-gather (NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors in the tests, with
the fill injected).  Shard boundaries are multiples of the K-b chunk (256 columns) when N allows, which keeps
every rank's first chunk free of run-in work; results do not depend on the boundaries either way."""
import os

import torch

CHUNK = 256


def column_shard(n_cols, world, rank, align=CHUNK):
    """[lo, hi) of rank `rank`: equal blocks of whole `align`-column chunks, remainder to the last ranks' left."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    units = -(-n_cols // align)                  # chunks, the last one possibly short
    base, extra = divmod(units, world)
    u_lo = rank * base + min(rank, extra)
    u_hi = u_lo + base + (1 if rank < extra else 0)
    return min(u_lo * align, n_cols), min(u_hi * align, n_cols)


def shard_table(n_cols, world, align=CHUNK):
    return [column_shard(n_cols, world, r, align) for r in range(world)]


class ShardedFill:
    """full[N cols, rows] (torch tensor, column-major matrix stored as rows of columns) filled by all ranks.

    fill(lo, hi, out_view) must write columns [lo, hi) into out_view (shape [hi-lo, rows]) and, when
    colstops is not None, their colstops into the int32 view it is given."""

    def __init__(self, n_cols, rows, device, dist=None, dtype=torch.float64):
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.n_cols, self.rows = n_cols, rows
        self.table = shard_table(n_cols, self.world)
        self.lo, self.hi = self.table[self.rank]
        self.full = torch.empty((n_cols, rows), dtype=dtype, device=device)
        self.colstops = torch.zeros(n_cols, dtype=torch.int32, device=device)
        self.equal = len({hi - lo for lo, hi in self.table}) == 1

    def local_view(self):
        return self.full[self.lo:self.hi], self.colstops[self.lo:self.hi]

    def fill_local(self, fill):
        out, cs = self.local_view()
        fill(self.lo, self.hi, out, cs)

    def gather(self):
        """reassemble on every rank (in place: each rank's shard already sits in its slice of `full`)"""
        if not self.dist:
            return self.full
        if self.equal:
            self.dist.all_gather_into_tensor(self.full.view(-1), self.full[self.lo:self.hi].reshape(-1))
            self.dist.all_gather_into_tensor(self.colstops, self.colstops[self.lo:self.hi].clone())
        else:  # ragged shard sizes: one broadcast per owner
            for r, (lo, hi) in enumerate(self.table):
                if hi > lo:
                    self.dist.broadcast(self.full[lo:hi], src=r)
                    self.dist.broadcast(self.colstops[lo:hi], src=r)
        return self.full

    def step(self, fill):
        self.fill_local(fill)
        return self.gather()


# ------------------------------------------------------------------------------------------ fused fill + all-gather
class _CudaView:
    """__cuda_array_interface__ holder: lets torch alias device memory owned by the C library"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class FusedShardedFill:
    """Every rank owns a full copy of the matrix in peer-visible memory; a rank fills ITS columns into EVERY copy
    from inside the transform kernel (pgb_transfer_fill_device_multi), so the shards are reassembled without a
    separate collective and only the retained rows of each chopped column cross NVLink.

    Memory: the library's own symmetric allocation (pgb_symm_*: cuMemCreate + fd export/import between the ranks, and an
    NVSwitch MULTICAST mapping where the fabric has one: one store per value reaches all ranks, the switch replicates
    it), else cudaMalloc + CUDA IPC (per-peer unicast stores).

    step(): zero the columns this rank does not own -> barrier -> fused fill -> barrier.  The barriers are
    stream-ordered (no host synchronisation): a one-warp kernel over peer-mapped flags (pgb_peer_barrier, the
    pre-fill one in split form inside the fill), or a 4-byte NCCL all-reduce when peer_barrier=False.  The whole
    step is graph-capturable."""

    def __init__(self, ctx, n_cols, rows, dist=None, peer_barrier=True, multicast=True, max_branches=64, max_nodes=None):
        import ctypes as C
        import numpy as np
        from . import _abi as A
        from . import engine as E
        self.C, self.A, self.ctx = C, A, ctx
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        if self.world > 8:
            raise ValueError("at most 8 peers (one NVSwitch domain)")
        self.n_cols, self.rows, self.ld = n_cols, rows, rows
        self.table = shard_table(n_cols, self.world)
        self.lo, self.hi = self.table[self.rank]
        dev = torch.device("cuda", ctx.device)
        self._tok = torch.zeros(1, dtype=torch.float32, device=dev)
        self._opened, self._clean = [], False
        self.peer_barrier = peer_barrier and self.dist is not None
        self.mc_mat = self.mc_cs = 0
        self.memory = "cudaMalloc + CUDA IPC"
        # one region: matrix | colstops | barrier flags (8 slots + this rank's epoch counter at slot 16)
        mat_bytes = 8 * n_cols * rows
        cs_off = (mat_bytes + 255) // 256 * 256
        fl_off = cs_off + (4 * max(n_cols, 1) + 255) // 256 * 256
        # scratch for the sharded K-a: (U, ULO, W, TWO, SN)[branches][nodes], published through the multicast mapping
        inv_off = fl_off + 256
        self.inv_cap = 5 * max_branches * (max_nodes or rows)
        total = inv_off + 8 * self.inv_cap
        base, peers, mc = None, None, 0
        self._symm = None
        if self.dist and not os.environ.get("PGB_NO_SYMM"):
            # symmetric memory owned by the library (pgb_symm_*: cuMemCreate + POSIX-fd export, pidfd_getfd import,
            # cuMulticast*): peer pointers for every rank's copy and, where the fabric has it, one multicast address
            def agree(flag):   # all ranks take the same path: a step counts only if it worked everywhere
                t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
                return t.item() == 1.0
            L = A.lib()
            S, own_fd, mc_fd = C.c_void_p(), C.c_int32(-1), C.c_int32(-1)
            ok = L.pgb_symm_begin(ctx.h, self.rank, self.world, total, 1 if multicast else 0, C.byref(S), C.byref(own_fd), C.byref(mc_fd)) == 0
            if agree(ok):
                info = [None] * self.world
                self.dist.all_gather_object(info, (os.getpid(), own_fd.value, mc_fd.value))
                pids = (C.c_int32 * self.world)(*[i[0] for i in info])
                fds = (C.c_int32 * self.world)(*[i[1] for i in info])
                ok = L.pgb_symm_connect(S, pids, fds, info[0][2]) == 0
                if agree(ok):               # every device has joined the multicast team: bind, then wait for all binds
                    ok = L.pgb_symm_bind(S) == 0
                    if agree(ok):
                        base = L.pgb_symm_ptr(S, self.rank)
                        peers = [int(L.pgb_symm_ptr(S, p)) for p in range(self.world)]
                        mc = int(L.pgb_symm_multicast_ptr(S) or 0)
                        self._symm = S
            if self._symm is None:
                if not ok:
                    import sys
                    msg = L.pgb_last_error(ctx.h)
                    sys.stderr.write(f"[pgb] rank {self.rank}: library symmetric memory unavailable ({msg.decode() if msg else '?'}); using CUDA IPC\n")
                if S:
                    L.pgb_symm_destroy(S)
        if base is None:
            self._buf = E.DeviceBuffer(ctx, total)                   # plain cudaMalloc: IPC-exportable
            base = self._buf.ptr.value
            peers = [None] * self.world
            peers[self.rank] = base
            if self.dist:
                h = (C.c_char * 64)()
                A.check(A.lib().pgb_ipc_export(ctx.h, self._buf.ptr, h), ctx.h)
                allh = [None] * self.world
                self.dist.all_gather_object(allh, bytes(h.raw))
                for r, blob in enumerate(allh):
                    if r == self.rank:
                        continue
                    p = C.c_void_p()
                    A.check(A.lib().pgb_ipc_open(ctx.h, (C.c_char * 64).from_buffer_copy(blob), C.byref(p)), ctx.h)
                    peers[r] = p.value
                    self._opened.append(p)
        else:
            self.memory = "library symmetric memory (cuMem*)" + (" + NVSwitch multicast (cuMulticast*)" if mc else ", per-peer stores")
        self.base = base
        self.peer_mat = list(peers)
        self.peer_cs = [p + cs_off for p in peers]
        self.peer_flags = [p + fl_off for p in peers]
        self.inv_sym = base + inv_off
        self.inv_mc = 0
        if mc:
            self.mc_mat, self.mc_cs = mc, mc + cs_off
            self.inv_mc = mc + inv_off
        self.full = torch.as_tensor(_CudaView(base, (n_cols, rows), "<f8"), device=dev)
        self.colstops = torch.as_tensor(_CudaView(base + cs_off, (n_cols,), "<i4"), device=dev)
        self._flags_t = torch.as_tensor(_CudaView(base + fl_off, (64,), "<i4"), device=dev)
        self.colstops.zero_()
        self._flags_t.zero_()
        # what every copy of the matrix is known to hold: zero from row prev[c] on in column c ("rows" = nothing known).
        # Local to the rank, but identical on all ranks: a fused fill writes every copy alike.
        self.prev = torch.full((n_cols,), rows, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        ctx.sync()
        if self.dist:
            self.dist.barrier()

    def barrier(self):
        """all ranks have finished what precedes this point on their streams (stream-ordered; no host sync)"""
        if not self.dist:
            return
        if self.peer_barrier:
            C, A = self.C, self.A
            fl = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in self.peer_flags])
            A.check(A.lib().pgb_peer_barrier(self.ctx.h, fl, self.world, self.rank, C.c_void_p(self.peer_flags[self.rank] + 4 * 16)), self.ctx.h)
        else:
            self.dist.all_reduce(self._tok)

    def zero_foreign(self, chopped=True):
        """clear the columns this rank does not own.  First time (or after an unchopped fill): the whole
        block; afterwards only what the peers wrote -- rows below the recorded colstops."""
        A, C = self.A, self.C
        spans = ((0, self.lo), (self.hi, self.n_cols))
        for a, b in spans:
            if b <= a:
                continue
            if self._clean:
                A.check(A.lib().pgb_zero_columns_ragged(self.ctx.h, C.c_void_p(self.base), self.ld, C.c_void_p(self.peer_cs[self.rank]), a, b), self.ctx.h)
            else:
                A.check(A.lib().pgb_zero_columns(self.ctx.h, C.c_void_p(self.base), self.ld, self.rows, a, b), self.ctx.h)
        self._clean = bool(chopped)

    def fill(self, dm, n, opts, with_barrier=False, sparse=False):
        """with_barrier: the pre-fill barrier is done by the fill in split form (arrive, K-a, K-b, wait, K-c);
        sparse: overwrite max(new, old colstop) rows of every copy (self.prev) -- nothing needs zeroing first"""
        A, C = self.A, self.C
        lo, hi = self.lo, self.hi
        outs = (C.c_void_p * self.world)(*[C.c_void_p(p + 8 * self.ld * lo) for p in self.peer_mat])
        css = (C.c_void_p * self.world)(*[C.c_void_p(p + 4 * lo) for p in self.peer_cs])
        if with_barrier:
            fl = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in self.peer_flags])
            ep = C.c_void_p(self.peer_flags[self.rank] + 4 * 16)
        else:
            fl, ep = C.cast(None, C.POINTER(C.c_void_p)), C.c_void_p()
        mco = C.c_void_p(self.mc_mat + 8 * self.ld * lo) if self.mc_mat else C.c_void_p()
        mcc = C.c_void_p(self.mc_cs + 4 * lo) if self.mc_mat else C.c_void_p()
        A.check(A.lib().pgb_transfer_fill_device_multi(self.ctx.h, dm.h, n, lo, hi, self.rows, outs, css, self.world, self.rank,
                                                       self.ld, C.byref(opts), fl, ep, mco, mcc,
                                                       C.c_void_p(self.prev.data_ptr() + 4 * lo) if sparse else C.c_void_p(),
                                                       C.c_void_p(self.inv_sym) if self.inv_mc else C.c_void_p(),
                                                       C.c_void_p(self.inv_mc) if self.inv_mc else C.c_void_p(), self.inv_cap), self.ctx.h)

    def step(self, dm, n, opts):
        """chopped output (the reference's): barrier ("my copy may be overwritten") -> fused fill storing
        max(new, old colstop) rows of every column to every rank -> barrier.  No zeroing pass: the copies keep the
        invariant "column c is zero from row prev[c] on" from one fill to the next.  Unchopped output: every rank
        first zeroes the columns it does not own."""
        sparse = bool(opts.chop)
        if not sparse:
            self.zero_foreign(chopped=bool(opts.chop))
            self.prev.fill_(self.rows)
        if self.peer_barrier:
            self.fill(dm, n, opts, with_barrier=True, sparse=sparse)
        else:
            self.barrier()
            self.fill(dm, n, opts, sparse=sparse)
        self.barrier()
        return self.full

    def close(self):
        self.ctx.sync()
        for p in self._opened:
            self.A.lib().pgb_ipc_close(self.ctx.h, p)
        self._opened = []
        if self._symm is not None:
            if self.dist:
                self.dist.barrier()
            self.A.lib().pgb_symm_destroy(self._symm)
            self._symm = None
