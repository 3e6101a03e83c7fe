"""Column sharding of the transfer-matrix fill across one-process-per-GPU ranks.

Mode columns are independent given the inverse branches, so rank r of W fills a contiguous block of columns
straight into its slice of the full column-major N x rows buffer and the shards are reassembled on every rank
with one all-gather (NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors in the tests, with
the fill injected).  Shard boundaries are multiples of the K-b chunk (256 columns) when N allows, which keeps
every rank's first chunk free of run-in work; results do not depend on the boundaries either way."""
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
