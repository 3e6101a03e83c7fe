"""Scratch: does an L2-sized K-b -> K-c staging slab beat the whole-block slab?  PGB_SLAB_COLS is read per fill."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases

E = P.engine
ctx = E.get_context()
n = N = 8192
m = cases.branches32()
dm = E.DeviceMap(m)
blk = E.DeviceBlock(ctx, n, N)
ref = None
for sc in [0, 4096, 2048, 1536, 1280, 1024, 768, 512, 256]:
    if sc: os.environ["PGB_SLAB_COLS"] = str(sc)
    else: os.environ.pop("PGB_SLAB_COLS", None)
    for _ in range(3):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk)
    ctx.sync()
    K = 20
    ctx.timer_start()
    for _ in range(K):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk)
    ms = ctx.timer_stop() / K
    ctx.profile(True)
    for _ in range(K):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk)
    prof = {k: ctx.profile_read(i) for i, k in enumerate(["invert", "basis", "transform"])}
    ctx.profile(False)
    a = blk.to_host()
    if ref is None: ref = a
    print(json.dumps({"slab_cols": sc, "ms": ms, "same_bits": bool(np.array_equal(a, ref)),
                      "prof_ms_per_fill": {k: v[0] / K for k, v in prof.items()}}), flush=True)
