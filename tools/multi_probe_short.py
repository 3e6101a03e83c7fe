"""Short variant of multi_probe.py: the fused multi-GPU step and its fill part only (run under torchrun), for A/B switches."""
import os, sys, json, ctypes as C
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import poltergeist_jl_b200 as P
import cases
from poltergeist_jl_b200.sharding import FusedShardedFill
E, A = P.engine, P._abi
local = int(os.environ.get("LOCAL_RANK", "0")); rank = int(os.environ.get("RANK", "0")); ws = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
ctx = E.get_context(local)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
N = n = 8192
dm = E.DeviceMap(cases.branches32(), ctx)
opts = A.pgb_chop_opts(0.0, 1, 0)
F = FusedShardedFill(ctx, N, n, dist)
def fill_multi(): dm.invalidate(); F.fill(dm, n, opts, sparse=True)
def step(): dm.invalidate(); F.step(dm, n, opts)
def timed(fn, reps=200):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    g = ctx.graph_capture(fn)
    run = lambda: ctx.graph_launch(g)
    for _ in range(3): run()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): run()
    e1.record(stream); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
step()
out = {"ws": ws, "PGB_MC_FLAGS": os.environ.get("PGB_MC_FLAGS"), "fill_multi_graph": timed(fill_multi), "step_graph": timed(step)}
if rank == 0: print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
