"""BASELINE configs[4]: batched linear-response sweep -- 1024 perturbations eps_i of the 2-branch base map,
N = n = 256 each: one batched fill (K-a, K-b, K-c over all maps) + batched QR + solves per sweep.
Under torchrun the maps are split across the ranks (replicas: no exchange except the gather of the acims)."""
import os, sys, json, time
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import poltergeist_jl_b200 as P
import cases
ws = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if ws > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nm, n = 1024, 256
epss = -0.05 + 0.1 * np.arange(nm) / (nm - 1)
mine = np.array_split(np.arange(nm), ws)[rank]
ctx = P.engine.get_context(local)
t0 = time.perf_counter()
S = P.BatchSweep([cases.perturbed(float(epss[i])) for i in mine], ctx)
setup_s = time.perf_counter() - t0
for _ in range(3):
    rho, _ = S.fill_acim(n, n)
if dist: dist.barrier()
torch.cuda.synchronize()
reps = 20
t0 = time.perf_counter()
for _ in range(reps):
    rho, _ = S.fill_acim(n, n)          # host call: fill + QR + solve + acims back on the host
    if dist:
        allr = [torch.empty((len(np.array_split(np.arange(nm), ws)[r]), n), dtype=torch.float64, device=f"cuda:{local}") for r in range(ws)]
        dist.all_gather(allr, torch.from_numpy(rho).to(f"cuda:{local}"))
if dist: dist.barrier()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
if dist:
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
if rank == 0:
    print(json.dumps({"config": "configs[4]: 1024 perturbations, N=n=256", "n_gpus": ws, "ms_per_sweep": dt * 1e3, "maps_per_s": nm / dt,
                      "entries_per_s": nm * n * n / dt, "descriptor_setup_s_host": setup_s}))
    if os.environ.get("PGB_TRACE"):
        pass
if dist:
    dist.barrier(); dist.destroy_process_group()
