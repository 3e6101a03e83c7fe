"""K-b time against the number of 256-column chunks of one launch (config 4, all 8192 nodes): intercept = per-launch fixed cost
(set-up of every node's tables and anchor states, pipeline fill/drain), slope = one chunk."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases
E = P.engine
ctx = E.get_context()
n = 8192
dm = E.DeviceMap(cases.branches32())
for lo, hi in [(0, 256), (0, 512), (0, 1024), (0, 2048), (0, 4096), (0, 8192), (7936, 8192), (7168, 8192)]:
    blk = E.DeviceBlock(ctx, n, hi - lo)
    blk.colstops.upload(np.full(hi - lo, n, dtype=np.int32))
    for _ in range(3):
        dm.invalidate(); E.transfer_fill_device(dm, n, lo, hi, out=blk, chop=True, sparse=True)
    ctx.sync()
    ctx.profile(True)
    K = 20
    for _ in range(K):
        dm.invalidate(); E.transfer_fill_device(dm, n, lo, hi, out=blk, chop=True, sparse=True)
    prof = {k: ctx.profile_read(i) for i, k in enumerate(["invert", "basis", "transform"])}
    ctx.profile(False)
    print(json.dumps({"cols": [lo, hi], "chunks": (hi - lo) // 256, "basis_ms": round(prof["basis"][0] / prof["basis"][1], 5),
                      "transform_ms": round(prof["transform"][0] / prof["transform"][1], 5)}))
    del blk
