"""Scratch: a fill cut into 512-column slabs (32 launches), eager vs replayed as ONE CUDA graph, with and without a
device->host copy saturating the link.  Does graph replay escape the per-launch penalty seen under D2H load?"""
import os, sys, json, time
os.environ["PGB_SLAB_COLS"] = os.environ.get("PGB_SLAB_COLS", "512")
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
src = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dst = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
side = torch.cuda.Stream()

def fill():
    dm.invalidate(); E.transfer_fill_device(dm, n, 0, N, out=blk, chop=True)

for _ in range(3): fill()
ctx.sync()
g = ctx.graph_capture(fill)
for _ in range(3): ctx.graph_launch(g)
ctx.sync()

def timed(fn, copying, K=8):
    torch.cuda.synchronize(); ctx.sync()
    t_enq = 0.0
    if copying:
        t0 = time.perf_counter()
        with torch.cuda.stream(side):
            for _ in range(4): dst.copy_(src, non_blocking=True)     # 4 x 256 MB = 18 ms of D2H in flight
        t_enq = time.perf_counter() - t0
    ctx.timer_start()
    for _ in range(K): fn()
    ms = ctx.timer_stop() / K
    torch.cuda.synchronize()
    return ms, t_enq * 1e3

res = {"slab_cols": os.environ["PGB_SLAB_COLS"]}
for name, fn in (("eager", fill), ("graph", lambda: ctx.graph_launch(g))):
    for copying in (False, True):
        ms, enq = timed(fn, copying)
        res[f"{name}_{'d2h' if copying else 'quiet'}_ms"] = round(ms, 4)
        if copying: res[f"{name}_copy_enqueue_ms"] = round(enq, 3)
print(json.dumps(res))
