"""Pinned-host <-> device copy bandwidth of this box (context for the e2e numbers)."""
import json, time, torch
dev = torch.device("cuda", 0)
out = {}
for mb in (64, 512):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for name, (dst, src) in {"d2h": (h, d), "h2d": (d, h)}.items():
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        out[f"{name}_{mb}MiB_GBs"] = 5 * n / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
