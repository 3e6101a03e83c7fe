"""Where the fused multi-GPU step spends its time (run under torchrun): CUDA-event times of its pieces."""
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
lo, hi = F.lo, F.hi
prev_local = torch.full((N,), n, dtype=torch.int32, device=dev)
def fill_local():
    dm.invalidate()
    cptr = C.c_void_p(prev_local.data_ptr() + 4 * lo)
    A.check(A.lib().pgb_transfer_fill_device_sparse(ctx.h, dm.h, n, lo, hi, n, C.c_void_p(F.base + 8 * n * lo), n, C.c_void_p(F.peer_cs[F.rank] + 4 * lo), cptr, C.byref(opts)), ctx.h)
def fill_multi():
    dm.invalidate(); F.fill(dm, n, opts, sparse=True)
def bar2(): F.barrier(); F.barrier()
def step(): dm.invalidate(); F.step(dm, n, opts)
def bar_fill_bar(): F.barrier(); fill_multi(); F.barrier()
def fill_split(): dm.invalidate(); F.fill(dm, n, opts, with_barrier=True, sparse=True)
def bar1(): F.barrier()
def ka_only():
    dm.invalidate()
    A.check(A.lib().pgb_transfer_fill_device_sparse(ctx.h, dm.h, n, lo, lo, n, C.c_void_p(F.base), n, C.c_void_p(F.peer_cs[F.rank]), C.c_void_p(prev_local.data_ptr()), C.byref(opts)), ctx.h)
def timed(fn, reps=100, graph=False):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    g = ctx.graph_capture(fn) if graph else None
    run = (lambda: ctx.graph_launch(g)) if g else fn
    for _ in range(3): run()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): run()
    e1.record(stream); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
out = {"ws": ws, "memory": F.memory}
step()
for name, fn, g in [("fill_local", fill_local, False), ("fill_local_graph", fill_local, True), ("fill_multi_graph", fill_multi, True),
                    ("barrier_x2_graph", bar2, True), ("bar_fill_bar_graph", bar_fill_bar, True), ("fill_split_barrier_graph", fill_split, True),
                    ("barrier_x1_graph", bar1, True), ("step_eager", step, False), ("step_graph", step, True)]:
    out[name] = timed(fn, graph=g)
ctx.profile(True)
for _ in range(10): fill_local()
prof = {k: ctx.profile_read(i) for i, k in enumerate(("invert_branches", "basis", "transform"))}
ctx.profile(False)
out["kernel_ms_fill_local"] = {k: v[0] / max(v[1], 1) for k, v in prof.items()}
ctx.profile(True)
for _ in range(10): fill_multi()
prof = {k: ctx.profile_read(i) for i, k in enumerate(("invert_branches", "basis", "transform"))}
ctx.profile(False)
out["kernel_ms_fill_multi"] = {k: v[0] / max(v[1], 1) for k, v in prof.items()}
t = torch.tensor([out["kernel_ms_fill_multi"]["transform"]], device=dev, dtype=torch.float64)
gl = [torch.zeros_like(t) for _ in range(ws)]
dist.all_gather(gl, t)
out["transform_ms_fill_multi_per_rank"] = [float(x.item()) for x in gl]
if rank == 0:
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
