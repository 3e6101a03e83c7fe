"""K-a / K-b / K-c times of ONE rank's column shard of config 4 on one GPU (what rank r of W runs in the fused multi-GPU step,
without the peer stores): shows what the run-in from the anchor group's start costs the high ranks.  PGB_KBM_SC32=1 for 32-chunk anchor groups (the default is 16)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases
E = P.engine
ctx = E.get_context()
n = N = 8192
m = cases.branches32()
dm = E.DeviceMap(m)
for W in (8, 4, 2):
    for r in sorted({0, W // 2, W - 1}):
        lo, hi = r * N // W, (r + 1) * N // W
        blk = E.DeviceBlock(ctx, n, hi - lo)
        blk.colstops.upload(np.full(hi - lo, n, dtype=np.int32))
        for _ in range(3):
            dm.invalidate(); E.transfer_fill_device(dm, n, lo, hi, out=blk, chop=True, sparse=True)
        ctx.sync()
        K = 20
        ctx.timer_start()
        for _ in range(K):
            dm.invalidate(); E.transfer_fill_device(dm, n, lo, hi, out=blk, chop=True, sparse=True)
        ms = ctx.timer_stop() / K
        ctx.profile(True)
        for _ in range(K):
            dm.invalidate(); E.transfer_fill_device(dm, n, lo, hi, out=blk, chop=True, sparse=True)
        prof = {k: ctx.profile_read(i) for i, k in enumerate(["invert", "basis", "transform"])}
        ctx.profile(False)
        print(json.dumps({"sc32": bool(os.environ.get("PGB_KBM_SC32")), "world": W, "rank": r, "cols": [lo, hi], "ms": round(ms, 5),
                          "kernel_ms": {k: round(v[0] / max(v[1], 1), 5) for k, v in prof.items()}}))
        del blk
