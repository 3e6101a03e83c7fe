"""Lane-level emulation of basis_mma_kernel (poltergeist.jl_b200/csrc/pgb_basis_mma.cuh) in numpy: the same fragment
maps, staging indices, swizzle and flush indices as the CUDA code, with mma.sync m8n8k4.f64 modelled from its PTX
fragment layout.  It checks the INDEX ALGEBRA of the kernel on the CPU (tests/test_kbm_layout.py); the arithmetic
itself is checked on the GPU against the oracle."""
import numpy as np

SC, TP, TBP = 16, 258, 17


def sslot(b, j):
    return b * 8 + (j ^ (((b & 3) << 1) | ((b >> 2) & 1)))


def mma_m8n8k4(C, a, b):
    """C: [32 lanes][2], a, b: [32 lanes].  lane = 4g + t holds A[g][t], B[t][g], C[g][2t + e]."""
    Am = np.zeros((8, 4)); Bm = np.zeros((4, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        Am[g, t] = a[lane]; Bm[t, g] = b[lane]
    D = Am @ Bm
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        C[lane, 0] += D[g, 2 * t]; C[lane, 1] += D[g, 2 * t + 1]


def rot(c, s, rc, rs):
    return c * rc - s * rs, s * rc + c * rs


def emulate_cta(theta, w, k_lo, k_hi, group):
    """theta, w: [ncomp][8 nodes].  Returns {k: values[8 nodes]} for the columns of anchor group `group` in [k_lo, k_hi)."""
    ncomp = theta.shape[0]
    nq = (ncomp + 3) >> 2
    out = {}
    tile = np.full((2, 8, TP), np.nan)
    par = 0
    c_first = (k_lo // (SC * 256) + group) * SC
    if c_first * 256 >= k_hi:
        return out
    # per-warp (node) state
    th = np.zeros((8, 32)); ww = np.zeros((8, 32))
    th[:, :ncomp] = theta.T; ww[:, :ncomp] = w.T
    tst = np.zeros((8, 32 * TBP, 2))
    for warp in range(8):
        for lane in range(32):
            for h in range(2):
                c, s = np.cos((8 * h + 0.5) * th[warp, lane]), np.sin((8 * h + 0.5) * th[warp, lane])
                tst[warp, lane * TBP + 8 * h] = (c, s)
                for r in range(1, 8):
                    c, s = rot(c, s, np.cos(th[warp, lane]), np.sin(th[warp, lane]))
                    tst[warp, lane * TBP + 8 * h + r] = (c, s)
    Tc = np.zeros((8, 32, 2, 8)); Ts = np.zeros((8, 32, 2, 8))
    for warp in range(8):
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for h in range(2):
                for q in range(8):
                    Tc[warp, lane, h, q], Ts[warp, lane, h, q] = tst[warp, (4 * q + t) * TBP + 8 * h + g]
    # set-up: owner lane b -> states of the first chunk (exact at step 0, rotations by 32 theta) -> fragment pick-up
    Sc = np.zeros((8, 32, 8)); Ss = np.zeros((8, 32, 8)); Rc = np.zeros((8, 32, 8)); Rs = np.zeros((8, 32, 8))
    for warp in range(8):
        stg = np.zeros((32, 8, 2))
        for lane in range(32):
            m0 = c_first * 256 + 15.5
            vc, vs = ww[warp, lane] * np.cos(m0 * th[warp, lane]), ww[warp, lane] * np.sin(m0 * th[warp, lane])
            for j in range(8):
                stg[lane, j] = (vc, vs)
                if j < 7:
                    vc, vs = rot(vc, vs, np.cos(32 * th[warp, lane]), np.sin(32 * th[warp, lane]))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for q in range(8):
                Sc[warp, lane, q], Ss[warp, lane, q] = stg[4 * q + t, g]
                Rc[warp, lane, q], Rs[warp, lane, q] = np.cos(256 * th[warp, 4 * q + t]), np.sin(256 * th[warp, 4 * q + t])
    pend = None
    for cc in range(SC):
        cstart = (c_first + cc) * 256
        if cstart >= k_hi:
            break
        if cc > 0:
            Sc, Ss = rot(Sc, Ss, Rc, Rs)          # every lane rotates its own 8 states
        if cstart + 256 <= k_lo:
            continue
        for warp in range(8):
            P = np.zeros((2, 32, 2)); Q = np.zeros((2, 32, 2))
            for q in range(8):
                if q < nq:
                    for h in range(2):
                        mma_m8n8k4(P[h], Sc[warp, :, q], Tc[warp, :, h, q])
                        mma_m8n8k4(Q[h], Ss[warp, :, q], Ts[warp, :, h, q])
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for h in range(2):
                    kp = (16 * g + 8 + 4 * h + t) ^ ((g & 1) << 2)
                    km = (16 * g + 7 - 4 * h - t) ^ ((g & 1) << 2)
                    tile[par, warp, 2 * kp] = P[h, lane, 0] - Q[h, lane, 0]
                    tile[par, warp, 2 * kp + 1] = P[h, lane, 1] - Q[h, lane, 1]
                    tile[par, warp, 2 * km] = P[h, lane, 1] + Q[h, lane, 1]
                    tile[par, warp, 2 * km + 1] = P[h, lane, 0] + Q[h, lane, 0]
        for tid in range(256):
            np_, cg = tid & 3, tid >> 2
            for it in range(4):
                col = cg + 64 * it
                kk = cstart + col
                if k_lo <= kk < k_hi:
                    idx = 2 * ((col >> 1) ^ (((col >> 5) & 1) << 2)) + (col & 1)
                    out.setdefault(kk, np.full(8, np.nan))
                    out[kk][2 * np_] = tile[par, 2 * np_, idx]
                    out[kk][2 * np_ + 1] = tile[par, 2 * np_ + 1, idx]
        tile[par] = np.nan
        par ^= 1
    return out


def bank_conflicts():
    """worst-case shared-memory conflict degree of the kernel's access patterns (8-byte bank pairs, 16 per 128-byte row)"""
    res = {}
    # staging writes (STS.128, quarter warps of 8 lanes): 16-byte slot = chunk index mod 8
    worst = 0
    for h in range(2):
        for sign in range(2):
            for qw in range(4):
                slots = []
                for lane in range(8 * qw, 8 * qw + 8):
                    g, t = lane >> 2, lane & 3
                    k = (16 * g + 8 + 4 * h + t) if sign == 0 else (16 * g + 7 - 4 * h - t)
                    slots.append((k ^ ((g & 1) << 2)) & 7)
                worst = max(worst, max(slots.count(s) for s in set(slots)))
    res["tile_write"] = worst
    # flush reads (LDS.64, half warps of 16 lanes): 8-byte bank = address mod 16
    worst = 0
    for w0 in range(0, 256, 16):
        for it in range(4):
            for odd in range(2):
                banks = []
                for tid in range(w0, w0 + 16):
                    np_, cg = tid & 3, tid >> 2
                    col = cg + 64 * it
                    idx = 2 * ((col >> 1) ^ (((col >> 5) & 1) << 2)) + (col & 1)
                    banks.append(((2 * np_ + odd) * TP + idx) & 15)
                worst = max(worst, max(banks.count(s) for s in set(banks)))
    res["tile_read"] = worst
    return res
