"""CPU check of the index algebra of the tensor-core K-b kernel (pgb_basis_mma.cuh): a lane-level numpy emulation with
the kernel's own fragment maps, staging indices, swizzle and flush indices must reproduce sum_b w_b cos(k theta_b) for
every column of arbitrary windows (run-in from the anchor, partial chunks, partial branch blocks), and the shared-memory
access patterns must be conflict-free."""
import numpy as np
import pytest

import kbm_emulate as K


@pytest.mark.parametrize("ncomp,k_lo,k_hi", [(32, 0, 512), (20, 300, 700), (32, 3072, 3372), (17, 2040, 2060), (16, 1792, 2304)])
def test_emulated_kernel_matches_direct_sum(ncomp, k_lo, k_hi):
    rng = np.random.default_rng(ncomp + k_lo)
    th = rng.uniform(0, np.pi, (ncomp, 8))
    w = rng.uniform(0, 0.1, (ncomp, 8))
    got = {}
    for G in range((k_hi - 1) // (K.SC * 256) - k_lo // (K.SC * 256) + 1):
        got.update(K.emulate_cta(th, w, k_lo, k_hi, G))
    assert sorted(got) == list(range(k_lo, k_hi))          # every requested column written exactly, nothing else
    thl, wl = th.astype(np.longdouble), w.astype(np.longdouble)
    for k, v in got.items():
        ref = (wl * np.cos(k * thl)).sum(0).astype(float)
        assert np.abs(v - ref).max() < 2e-12                 # numpy's own cos(k theta) noise; index errors are O(1)


def test_shared_memory_patterns_are_conflict_free():
    assert K.bank_conflicts() == {"tile_write": 1, "tile_read": 1}
