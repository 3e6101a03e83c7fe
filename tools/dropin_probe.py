"""Latency of the drop-in path (Transfer -> resizedata! (adaptive) -> SolutionInv -> acim) per phase, with kernel-launch counts."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases
ctx = P.engine.get_context()
for name, ncols in (("readme_map", 128), ("lanford", 256), ("lanford", 1024), ("circle4", 4096), ("branches32", 8192)):
    m = getattr(cases, name)()
    best = None
    for rep in range(5):
        ctx.sync()
        l0 = ctx.launch_count(); t0 = time.perf_counter()
        L = P.Transfer(m)
        t1 = time.perf_counter()
        L.resizedata(ncols)
        ctx.sync()
        t2 = time.perf_counter(); l2 = ctx.launch_count()
        K = P.SolutionInv(L, n=ncols)
        t3 = time.perf_counter(); l3 = ctx.launch_count()
        rho = K.acim().coefficients
        t4 = time.perf_counter(); l4 = ctx.launch_count()
        rec = {"case": name, "columns": ncols, "create_ms": 1e3 * (t1 - t0), "resize_ms": 1e3 * (t2 - t1), "solinv_ms": 1e3 * (t3 - t2), "acim_ms": 1e3 * (t4 - t3),
               "total_ms": 1e3 * (t4 - t0), "launches": {"resize": l2 - l0, "solinv": l3 - l2, "acim": l4 - l3}, "kq": None,
               "max_nodes": int(np.max(L.nodes_used(0, ncols)))}
        K.close(); L.close()
        if best is None or rec["total_ms"] < best["total_ms"]:
            best = rec
    print(json.dumps(best))
