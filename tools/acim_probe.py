"""Phase times of acim at config 4 (PGB_TRACE=1 prints the C-side phases)."""
import os, sys, time, json
if "--trace" in sys.argv:
    os.environ["PGB_TRACE"] = "1"
    sys.argv.remove("--trace")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases
E = P.engine
ctx = E.get_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dm = E.DeviceMap(cases.branches32())
blk = E.transfer_fill_device(dm, n, 0, n, chop=True)
ctx.sync()
for structured in (True, True, True, True, False, False):
    t0 = time.perf_counter()
    K = E.DenseSolutionInv(blk, n, P.CHEBYSHEV, 0.0, 1.0, structured=structured)
    t1 = time.perf_counter()
    rho = K.acim()
    t2 = time.perf_counter()
    K.close()
    t3 = time.perf_counter()
    print(json.dumps({"structured": structured, "kq": None, "create_ms": (t1 - t0) * 1e3, "acim_ms": (t2 - t1) * 1e3, "close_ms": (t3 - t2) * 1e3}), flush=True)
