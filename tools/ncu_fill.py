"""Short program for ncu: 4 full config-4 fills (K-a + K-b + K-c) on one GPU; the first one writes every row of the block,
the others are the bench's steady state (K-c stores only rows below the colstops)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import poltergeist_jl_b200 as P
import cases
E = P.engine
ctx = E.get_context()
name = sys.argv[1] if len(sys.argv) > 1 else "branches32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dm = E.DeviceMap(getattr(cases, name)())
import numpy as np
blk = E.DeviceBlock(ctx, n, n)
blk.colstops.upload(np.full(n, n, dtype=np.int32))            # "nothing known about the block": the first sparse fill writes every row
for _ in range(4):
    dm.invalidate()
    E.transfer_fill_device(dm, n, 0, n, out=blk, chop=True, sparse=True)   # the bench's configuration: chopped output, sparse store
ctx.sync()
print("ok", ctx.launch_count())
