"""Scratch timing of the fixed-grid fill (device resident) with the per-kernel profile."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases
CHOP = os.environ.get("QT_CHOP", "0") == "1"
SPARSE = CHOP and os.environ.get("QT_SPARSE", "0") == "1"   # pgb_transfer_fill_device_sparse: K-c writes max(new, old colstop) rows

E = P.engine
ctx = E.get_context()
print("fp64 peak TF", ctx.fp64_peak_tflops())
for name, n, N in [("branches32", 8192, 8192), ("circle4", 4096, 4096), ("lanford", 1024, 1024), ("readme_map", 8192, 8192)]:
    m = getattr(cases, name)()
    dm = E.DeviceMap(m)
    blk = E.DeviceBlock(ctx, n, N)
    if SPARSE:
        blk.colstops.upload(np.full(N, n, dtype=np.int32))
    for _ in range(3):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk, chop=CHOP, sparse=SPARSE)
    ctx.sync()
    K = 10
    ctx.timer_start()
    for _ in range(K):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk, chop=CHOP, sparse=SPARSE)
    ms = ctx.timer_stop() / K
    ctx.profile(True)
    for _ in range(K):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk, chop=CHOP, sparse=SPARSE)
    prof = {k: ctx.profile_read(i) for i, k in enumerate(["invert", "basis", "transform"])}
    ctx.profile(False)
    B = dm.ncomposite
    print(json.dumps({"chop": CHOP, "sparse": SPARSE, "kb_legacy": bool(os.environ.get("PGB_KB_LEGACY")), "case": name, "n": n, "N": N, "B": B, "ms": ms, "entries_per_s": n * N / ms * 1e3,
                      "prof_ms_per_launch": {k: v[0] / max(v[1], 1) for k, v in prof.items()},
                      "basis_TF": 4.0 * B * n * N / (prof["basis"][0] / prof["basis"][1] * 1e-3) / 1e12,
                      "transform_GBs": 16.0 * n * N / (prof["transform"][0] / prof["transform"][1] * 1e-3) / 1e9}))
