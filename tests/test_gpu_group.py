"""pgb_group: several GPUs driven from ONE process with library-owned symmetric memory (cuMem* / cuMulticast*): the fused
sharded fill and the single-host-buffer RaggedMatrix call against the single-GPU results, bit for bit (columns do not
depend on how a request is cut).  With one visible GPU the group has one member (same code path, no peers); the
multi-device cases run when the box has more (gpurun --gpus 2)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("multicast", [True, False])
def test_group_fill_matches_single_gpu(P, world, multicast):
    if world > _ndev():
        pytest.skip(f"needs {world} GPUs")
    E = P.engine
    m = cases.branches32()
    n = N = 2048
    want, wcs = E.transfer_fill(m, n, 0, N, chop=True)
    G = E.Group(list(range(world)))
    gm = G.map(m)
    M = G.matrix(n, N, multicast=multicast)
    assert not (M.multicast and not multicast)        # (multicast may be unavailable on a fabric: the unicast path is still correct)
    print(f"\ngroup of {world}: multicast requested {multicast}, in use {M.multicast}")
    slow = G.map(P.modulomap(P.PolySine(p1=12.0, A=3.0 / (2 * np.pi), w=2 * np.pi), P.Interval(0.0, 1.0)))
    for which in (slow, gm, gm):                      # longer columns first: the next fill must clear their tails in EVERY copy
        G.fill(which, n, M, chop=True)
        G.sync()
    for r in range(world):
        assert np.array_equal(M.colstops_host(r), wcs), r
        assert np.array_equal(M.to_host(r), want), r
    # the solves run on one device, on its copy
    K = E.DenseSolutionInv(M.block(0), N, P.CHEBYSHEV, 0.0, 1.0, structured=True)
    rho = K.acim()
    assert abs(rho[0] - 1.0) < 1e-12 and np.abs(rho[1:]).max() < 1e-12
    K.close()
    # raw (unchopped) fill through the same path
    G.fill(gm, n, M, chop=False)
    G.sync()
    raw, _ = E.transfer_fill(m, n, 0, N, chop=False)
    for r in range(world):
        assert np.array_equal(M.to_host(r), raw), r
    M.close(); G.close()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_group_fill_ragged_is_one_ragged_matrix(P, world):
    if world > _ndev():
        pytest.skip(f"needs {world} GPUs")
    E = P.engine
    G = E.Group(list(range(world)))
    for m, n, lo, hi in ((cases.branches32(), 2048, 0, 2048), (cases.branches32(), 1024, 100, 900), (cases.lanford(), 256, 0, 200),
                         (cases.circle4(), 512, 3, 4)):
        d0, c0, s0, m0 = E.transfer_fill_ragged(m, n, lo, hi)
        gm = G.map(m)
        d1, c1, s1, m1 = G.fill_ragged(gm, n, lo, hi)
        assert m0 == m1 and np.array_equal(c0, c1) and np.array_equal(s0, s1)
        assert np.array_equal(d0, d1)
    G.close()


def test_symmetric_memory_single_process_roundtrip(P):
    """every copy is addressable from every device, and a store through the multicast address reaches all copies"""
    world = min(_ndev(), 2)
    E = P.engine
    G = E.Group(list(range(world)))
    M = G.matrix(64, 16)
    m = cases.lanford()
    G.fill(G.map(m), 64, M, chop=True)
    G.sync()
    a = M.to_host(0)
    for r in range(1, world):
        assert np.array_equal(M.to_host(r), a)
    assert np.abs(a).max() > 0.1
    M.close(); G.close()
