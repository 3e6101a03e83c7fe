"""CPU tests of the N > 1 host logic: column sharding + reassembly over torch.distributed (gloo, world size 2,
127.0.0.1).  The CUDA fill is replaced by an injected fill (the oracle's exact block -- test infrastructure --
or a synthetic function of the column index); what is under test is the shard table, the in-place all-gather and
the ragged-shard fallback, i.e. the code bench.py and a multi-GPU host run with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _synthetic(lo, hi, out, cs):
    cols = torch.arange(lo, hi, dtype=torch.float64).unsqueeze(1)
    rows = torch.arange(out.shape[1], dtype=torch.float64).unsqueeze(0)
    out.copy_(torch.sin(0.37 * cols + 0.11 * rows) + cols)
    cs.copy_(torch.arange(lo, hi, dtype=torch.int32) % 17)


def _worker(rank, world, port, n_cols, rows, use_oracle, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import poltergeist_jl_b200 as P
    from poltergeist_jl_b200.sharding import ShardedFill
    S = ShardedFill(n_cols, rows, torch.device("cpu"), dist)
    if use_oracle:
        from oracle import oracle as O
        desc = cases.lanford().to_desc()

        def fill(lo, hi, out, cs):
            if hi > lo:
                out.copy_(torch.from_numpy(np.ascontiguousarray(O.exact_dense(desc, rows, lo, hi, rows).T)))
                cs.fill_(rows)
    else:
        fill = _synthetic
    full = S.step(fill)
    dist.barrier()
    if rank == 0:
        q.put((full.numpy().copy(), S.colstops.numpy().copy(), S.table))
    dist.destroy_process_group()


def _run(n_cols, rows, use_oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cols, rows, use_oracle, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_shard_table():
    from poltergeist_jl_b200.sharding import column_shard, shard_table
    assert shard_table(8192, 8) == [(1024 * r, 1024 * (r + 1)) for r in range(8)]
    assert shard_table(8192, 1) == [(0, 8192)]
    for n, w in ((8192, 2), (8192, 4), (1000, 3), (256, 8), (100, 2), (4097, 8)):
        t = shard_table(n, w)
        assert t[0][0] == 0 and t[-1][1] == n
        assert all(t[i][1] == t[i + 1][0] for i in range(w - 1))          # contiguous cover, no overlap
        assert all(lo % 256 == 0 for lo, hi in t if hi > lo)               # shards start on K-b chunk boundaries
    with pytest.raises(ValueError):
        column_shard(10, 2, 2)


def test_two_rank_gather_equal_shards():
    full, cs, table = _run(512, 32, False)
    want = torch.empty((512, 32), dtype=torch.float64)
    wcs = torch.empty(512, dtype=torch.int32)
    _synthetic(0, 512, want, wcs)
    assert table == [(0, 256), (256, 512)]
    assert np.array_equal(full, want.numpy()) and np.array_equal(cs, wcs.numpy())


def test_two_rank_gather_ragged_shards():
    full, cs, table = _run(300, 16, False)        # 2 chunks: [0,256) and [256,300)
    want = torch.empty((300, 16), dtype=torch.float64)
    wcs = torch.empty(300, dtype=torch.int32)
    _synthetic(0, 300, want, wcs)
    assert table == [(0, 256), (256, 300)]
    assert np.array_equal(full, want.numpy()) and np.array_equal(cs, wcs.numpy())


def test_two_rank_sharded_matrix_equals_unsharded(O):
    """the reassembled sharded matrix is the unsharded one (oracle block as the injected fill)"""
    full, cs, table = _run(512, 64, True)
    E = O.exact_dense(cases.lanford().to_desc(), 64, 0, 512, 64)
    assert np.array_equal(full, np.ascontiguousarray(E.T))


def test_shard_table_properties():
    """property test: any (columns, world) gives a contiguous, chunk-aligned, balanced cover"""
    from hypothesis import given, settings, strategies as st
    from poltergeist_jl_b200.sharding import shard_table

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 20000), st.integers(1, 8))
    def check(n, w):
        t = shard_table(n, w)
        assert len(t) == w and t[0][0] == 0 and t[-1][1] == n
        sizes = []
        for i, (lo, hi) in enumerate(t):
            assert 0 <= lo <= hi <= n
            if i + 1 < w:
                assert hi == t[i + 1][0]
            if hi > lo:
                assert lo % 256 == 0
                sizes.append(hi - lo)
        chunks = [-(-z // 256) for z in sizes]
        assert max(chunks) - min(chunks) <= 1                             # whole chunks, spread evenly
    check()
