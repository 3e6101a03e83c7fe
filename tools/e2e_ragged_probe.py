"""e2e of pgb_transfer_fill_ragged (config 4) -- run with PGB_RAGGED_SLABS=k to see the pipeline depth's effect."""
import os, sys, json, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import poltergeist_jl_b200 as P
import cases
E, A = P.engine, P._abi
ctx = E.get_context()
n = 8192
m = cases.branches32(); desc = m.to_desc()
opts = A.pgb_chop_opts(0.0, 1, 0)
hdata = torch.empty(n * n // 4, dtype=torch.float64).pin_memory()
hcols = torch.empty(n + 1, dtype=torch.int64).pin_memory(); hcs = torch.empty(n, dtype=torch.int32).pin_memory()
mm, nnz, off = C.c_int64(0), C.c_int64(0), C.c_int64(0)
TAIL = os.environ.get("RAGGED_TAIL", "0") == "1"
dmp = E.DeviceMap(m, ctx)
def step(newmap=True):
    if newmap:
        mh = C.c_void_p(); A.check(A.lib().pgb_map_create(ctx.h, C.byref(desc), C.byref(mh)), ctx.h)
    else:
        dmp.invalidate(); mh = dmp.h
    if TAIL:
        A.check(A.lib().pgb_transfer_fill_ragged_tail(ctx.h, mh, n, 0, n, C.cast(hdata.data_ptr(), C.POINTER(C.c_double)), hdata.numel(), C.byref(off),
                C.cast(hcols.data_ptr(), C.POINTER(C.c_int64)), C.cast(hcs.data_ptr(), C.POINTER(C.c_int32)), C.byref(mm), C.byref(nnz), C.byref(opts)), ctx.h)
    else:
        A.check(A.lib().pgb_transfer_fill_ragged(ctx.h, mh, n, 0, n, C.cast(hdata.data_ptr(), C.POINTER(C.c_double)), hdata.numel(),
                C.cast(hcols.data_ptr(), C.POINTER(C.c_int64)), C.cast(hcs.data_ptr(), C.POINTER(C.c_int32)), C.byref(mm), C.byref(nnz), C.byref(opts)), ctx.h)
    if newmap: A.lib().pgb_map_destroy(mh)
out = {"tail": TAIL, "fracs": os.environ.get("PGB_RAGGED_FRACS", "default")}
for name, nm in (("new_map", True), ("same_map", False)):
    for _ in range(5): step(nm)
    t0 = time.perf_counter()
    for _ in range(50): step(nm)
    out[name + "_ms"] = (time.perf_counter() - t0) / 50 * 1e3
ctx.profile(True)
for _ in range(20): step(False)
ctx.sync()
out["kernel_ms_per_call"] = {k: ctx.profile_read(i)[0] / 20 for i, k in enumerate(["invert", "basis", "transform"])}
ctx.profile(False)
print(json.dumps(out))
