"""Does torch symmetric memory (peer pointers + NVSwitch multicast) work on this box?  (run under torchrun)"""
import os, sys, json
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as sm
out = {"rank": rank}
try:
    t = sm.empty((1 << 20,), dtype=torch.float64, device=torch.device("cuda", local))
    h = sm.rendezvous(t, dist.group.WORLD)
    out.update(ptrs=[hex(p) for p in h.buffer_ptrs], mc=hex(h.multicast_ptr),
               data_ptr=hex(t.data_ptr()), world=h.world_size)
except Exception as e:
    out["error"] = repr(e)[:500]
print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
