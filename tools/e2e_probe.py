"""Where the host-buffer call spends its time: map upload, fill, device->host copy (config 4)."""
import os, sys, json, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import poltergeist_jl_b200 as P
import cases
E, A = P.engine, P._abi
ctx = E.get_context()
n = 8192
m = cases.branches32()
desc = m.to_desc()
host = torch.empty((n, n), dtype=torch.float64).pin_memory()
hcs = np.zeros(n, dtype=np.int32)
opts = A.pgb_chop_opts(0.0, 0, 0)
def T(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
out = {}
mh = C.c_void_p()
def create_destroy():
    A.check(A.lib().pgb_map_create(ctx.h, C.byref(desc), C.byref(mh)), ctx.h); A.lib().pgb_map_destroy(mh)
out["map_create_destroy_ms"] = T(create_destroy)
dm = E.DeviceMap(m, ctx)
blk = E.DeviceBlock(ctx, n, n)
def fill_dev():
    dm.invalidate(); E.transfer_fill_device(dm, n, 0, n, out=blk); ctx.sync()
out["fill_device_sync_ms"] = T(fill_dev)
def d2h():
    A.check(A.lib().pgb_dev_download(ctx.h, C.c_void_p(host.data_ptr()), blk.buf.ptr, 8 * n * n), ctx.h)
out["d2h_512MiB_pinned_ms"] = T(d2h)
pageable = np.empty((n, n))
def d2h_pageable():
    A.check(A.lib().pgb_dev_download(ctx.h, pageable.ctypes.data_as(C.c_void_p), blk.buf.ptr, 8 * n * n), ctx.h)
out["d2h_512MiB_pageable_ms"] = T(d2h_pageable, 2)
def fill_host():
    dm.invalidate()
    A.check(A.lib().pgb_transfer_fill(ctx.h, dm.h, n, 0, n, n, C.cast(host.data_ptr(), C.POINTER(C.c_double)), n,
                                      hcs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(opts)), ctx.h)
out["fill_host_pinned_ms"] = T(fill_host)
def fill_ragged():
    dm.invalidate()
    return E.transfer_fill_ragged(dm, n, 0, n)
out["fill_ragged_ms"] = T(fill_ragged)
d, cols, cs, mm = fill_ragged()
out["ragged_nnz"] = int(len(d)); out["ragged_max_colstop"] = int(mm)
print(json.dumps(out))
