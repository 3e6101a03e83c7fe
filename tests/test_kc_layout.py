"""CPU check of the index scheme behind K-c's last pass at n = 8192 (csrc/pgb_kernels.cuh, `SpecRegs` and the LOGM == 12 branch of
transform_kernel): after the third radix-16 pass thread t holds W[t + 256 rr], rr = 0..15, in register r[bitrev4(rr)]; the
real-sequence split needs the pairs (W_j, W_{M-j}), j = 1..M/2-1, plus W_0 and W_{M/2}.  Thread t takes j = t + 256 it (it < 8)
from its own registers and W_{M-j} from the exchange area xch[(rr - 8) * 256 + t'], written by every thread for rr >= 8."""
import numpy as np

M, T = 4096, 256


def brev4(q):
    return ((q & 1) << 3) | ((q & 2) << 1) | ((q & 4) >> 1) | ((q & 8) >> 3)


def test_every_pair_is_found_exactly_once():
    W = np.arange(M)                                    # W[f] = f: values identify frequencies
    regs = np.empty((T, 16), dtype=np.int64)            # regs[t][K] = W[t + 256 * brev4(K)] (dft_regs leaves y[bitrev(K)] in r[K])
    for t in range(T):
        for K in range(16):
            regs[t, K] = W[t + 256 * brev4(K)]
    xch = np.full(M // 2, -1, dtype=np.int64)
    for t in range(T):
        for K in range(16):
            rr = brev4(K)
            if rr >= 8:
                assert xch[(rr - 8) * 256 + t] == -1
                xch[(rr - 8) * 256 + t] = regs[t, K]
    assert (xch >= 0).all()                             # the exchange area is written exactly once everywhere
    seen = set()
    for t in range(T):
        for it in range(8):
            j = t + 256 * it
            if j == 0:
                continue                                 # thread 0, slot 0: W_0 and W_{M/2} are the special values
            wj = regs[t, brev4(it)]
            wm = xch[(256 - t) + 256 * (7 - it)]
            assert wj == j and wm == M - j, (t, it)
            assert j not in seen
            seen.add(j)
    assert seen == set(range(1, M // 2))
    assert regs[0, 0] == 0 and regs[0, 1] == M // 2     # SpecRegs::w0(), wh()


def test_padded_addresses_decompose():
    """padc(g + 16 m) = padc(g) + 17 m: what lets every shared-memory address be a per-thread base plus a constant"""
    padc = lambda i: i + (i >> 4)
    for g in range(0, 4096, 7):
        for m in range(0, 256, 5):
            if g + 16 * m < 4096:
                assert padc(g + 16 * m) == padc(g) + 17 * m
    # store side of a Stockham pass: ob = q + s R p, q < s, index ob + s rr (R = 16)
    for s in (1, 16, 256):
        pstep = s + s // 16 if s % 16 == 0 else s
        for g in range(256):
            p, q = divmod(g, s)
            ob = q + s * 16 * p
            for rr in range(16):
                if ob + s * rr < 4096:
                    assert padc(ob + s * rr) == padc(ob) + pstep * rr
