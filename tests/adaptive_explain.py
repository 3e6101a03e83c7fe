"""TEST HELPER.  The reference's accept / chop decisions (Transfer.jl:137-150) evaluated in numpy on a raw coefficient
column, with the MARGIN of every decision -- so that a test can demand, for each column where the GPU path and the faithful
restatement decided differently, that the deciding quantity lies within the two implementations' entry deviation of
its threshold (a legitimate flip), instead of accepting a percentage of unexplained mismatches."""
import numpy as np

EPS = float(np.finfo(float).eps)
TOL = 200 * EPS


def checkpoints(O, desc):
    ra, rb = desc.ran_a, desc.ran_b
    if desc.range_space == O.CHEBYSHEV:
        return [(ra + rb) / 2 + (rb - ra) / 2 * c for c in (-0.823972, 0.01, 0.3273484)]
    return [(ra + rb) / 2 + (rb - ra) / 2 * (c / np.pi - 1.0) for c in (1.223972, 3.14, 5.83273484)]


def decisions(O, desc, raw, k):
    """raw: the n raw coefficients of column k.  -> dict with the accept test's quantities and thresholds"""
    n = len(raw)
    logn = int(round(np.log2(n)))
    mxc = max(np.abs(raw).max(), 1.0)
    rs = desc.range_space
    b = n if rs == O.CHEBYSHEV else n // 2 + 1
    b23 = max((2 * b) // 3, 1)
    bs = b23 if rs == O.CHEBYSHEV else (1 if b23 == 1 else 2 * b23 - 2)
    tail = np.abs(raw[bs - 1:]).max()
    tail_thr = 20 * TOL * mxc * logn
    r = checkpoints(O, desc)
    fr = np.array([O.ref_transferfunction(desc, x, k) for x in r])
    mfr = max(np.abs(fr).max(), 1.0)
    res = np.abs(O.fun_eval(rs, desc.ran_a, desc.ran_b, raw, np.array(r)) - fr)
    pt_thr = TOL * n * mfr * 500
    chop_thr = TOL * mxc * logn
    nz = np.nonzero(np.abs(raw) > chop_thr)[0]
    ln = int(nz[-1]) + 1 if len(nz) else 0
    accept = n > 8 and tail < tail_thr and bool(np.all(res < pt_thr))
    return {"accept": accept, "tail": tail, "tail_thr": tail_thr, "res": res, "pt_thr": pt_thr, "chop_thr": chop_thr, "len": ln}


def explain_column(P, O, m, desc, k, n_g, cs_g, n_r, cs_r):
    """Why do GPU (grid n_g, colstop cs_g) and restatement (n_r, cs_r) differ on column k?  -> (explained, text)"""
    notes = []
    ok = True
    n_lo = min(n_g, n_r)
    if n_g != n_r:
        # one of them accepted the column on n_lo, the other did not: some accept criterion sits at its threshold
        G = P.engine.transfer_fill(m, n_lo, k, k + 1, chop=False)[0][:, 0]
        R = O.ref_dense(desc, n_lo, k, k + 1, n_lo)[:, 0]
        dev = np.abs(G - R).max()
        dg, dr = decisions(O, desc, G, k), decisions(O, desc, R, k)
        near = abs(dr["tail"] - dr["tail_thr"]) <= dev or abs(dg["tail"] - dg["tail_thr"]) <= dev
        # a checkpoint value is a sum of n coefficients: it moves by at most the 1-norm of the coefficient deviation
        dev1 = np.abs(G - R).sum() + 64 * EPS * max(np.abs(R).max(), 1.0)
        near = near or bool(np.any(np.abs(dr["res"] - dr["pt_thr"]) <= dev1)) or bool(np.any(np.abs(dg["res"] - dg["pt_thr"]) <= dev1))
        flipped = dg["accept"] != dr["accept"]
        notes.append(f"grid {n_g} vs {n_r}: on n={n_lo} gpu accept={dg['accept']} ref accept={dr['accept']} tail {dr['tail']:.3e}/{dr['tail_thr']:.3e} "
                     f"dev {dev:.1e} near={near}")
        ok = ok and flipped and near
    else:
        G = P.engine.transfer_fill(m, n_g, k, k + 1, chop=False)[0][:, 0]
        R = O.ref_dense(desc, n_g, k, k + 1, n_g)[:, 0]
        dev = np.abs(G - R).max()
        thr = decisions(O, desc, R, k)["chop_thr"]
        row = max(cs_g, cs_r) - 1
        margin = min(abs(abs(R[row]) - thr), abs(abs(G[row]) - thr))
        notes.append(f"colstop {cs_g} vs {cs_r} on n={n_g}: |c[{row}]| = {abs(R[row]):.4e}, threshold {thr:.4e}, margin {margin:.1e}, dev {dev:.1e}")
        ok = ok and margin <= dev + 4 * EPS
    return ok, "; ".join(notes)
