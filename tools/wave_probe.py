"""Scratch: K-b / K-c time against the number of 256-column chunks (wave quantisation of the K-b grid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import poltergeist_jl_b200 as P
import cases

E = P.engine
ctx = E.get_context()
n = 8192
dm = E.DeviceMap(cases.branches32())
for chunks in [8, 16, 20, 23, 24, 26, 27, 28, 29, 30, 31, 32]:
    N = 256 * chunks
    blk = E.DeviceBlock(ctx, n, N)
    for _ in range(3):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk)
    ctx.sync()
    K = 20
    ctx.profile(True)
    for _ in range(K):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk)
    prof = {k: ctx.profile_read(i) for i, k in enumerate(["invert", "basis", "transform"])}
    ctx.profile(False)
    b, t = prof["basis"][0] / K, prof["transform"][0] / K
    print(json.dumps({"chunks": chunks, "ctas": 256 * chunks, "waves_5": 256 * chunks / 740, "basis_ms": b, "basis_us_per_chunk": 1e3 * b / chunks,
                      "transform_ms": t, "transform_us_per_chunk": 1e3 * t / chunks}), flush=True)
    del blk
