"""Scratch: does a device->host DMA running beside the fill slow the fill's kernels?"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import poltergeist_jl_b200 as P
import cases

E = P.engine
ctx = E.get_context()
n = N = 8192
dm = E.DeviceMap(cases.branches32())
blk = E.DeviceBlock(ctx, n, N)
src = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
dst = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
side = torch.cuda.Stream()

def run(copying, K=20):
    for _ in range(3):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk, chop=True)
    ctx.sync(); torch.cuda.synchronize()
    ctx.profile(True)
    for _ in range(K):
        if copying:
            with torch.cuda.stream(side):
                for _ in range(2): dst.copy_(src, non_blocking=True)   # 2 x 64 MB = 2.3 ms of DMA behind every fill
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk, chop=True)
    ctx.sync(); torch.cuda.synchronize()
    prof = {k: ctx.profile_read(i) for i, k in enumerate(["invert", "basis", "transform"])}
    ctx.profile(False)
    return {k: v[0] / K for k, v in prof.items()}

def wall(copying, K=20):
    """device time of K fills by the context timer (includes launch gaps, unlike the per-kernel profile)"""
    for _ in range(3):
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk, chop=True)
    ctx.sync(); torch.cuda.synchronize()
    ctx.timer_start()
    for _ in range(K):
        if copying:
            with torch.cuda.stream(side):
                for _ in range(2): dst.copy_(src, non_blocking=True)
        dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk, chop=True)
    ms = ctx.timer_stop() / K
    torch.cuda.synchronize()
    return ms

res = {"slab_cols": os.environ.get("PGB_SLAB_COLS", "whole"), "no_copy": run(False), "with_d2h_copy": run(True),
       "wall_no_copy_ms": wall(False), "wall_with_d2h_copy_ms": wall(True)}
print(json.dumps(res))
