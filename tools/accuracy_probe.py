"""Entry-level accuracy of the GPU fill against the extended-precision oracle, beside the faithful
FP64 restatement's own error (|ref - exact|), in units of the column coefficient 2-norm."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases
from oracle import oracle as O

CASES = [("lanford", 1024, 0, 32), ("lanford", 1024, 480, 512), ("lanford", 1024, 992, 1024),
         ("branches32", 1024, 0, 16), ("branches32", 1024, 1008, 1024), ("circle4", 4096, 4064, 4096),
         ("readme_map", 2048, 2032, 2048), ("branches32", 8192, 4096, 4104), ("branches32", 8192, 8184, 8192)]
for name, n, lo, hi in CASES:
    m = getattr(cases, name)()
    d = m.to_desc()
    G, _ = P.engine.transfer_fill(m, n, lo, hi, chop=False)
    E = O.exact_dense(d, n, lo, hi, n, prec=0)
    R = O.ref_dense(d, n, lo, hi, n) if n <= 4096 else None
    nrm = np.linalg.norm(E, axis=0)
    floor = max(nrm.max() * 1e-3, 1e-300)
    den = np.maximum(nrm, floor)
    out = {"case": name, "n": n, "cols": [lo, hi], "gpu_vs_exact": float((np.abs(G - E).max(axis=0) / den).max()),
           "colnorm": [float(nrm.min()), float(nrm.max())]}
    if R is not None:
        out["ref_vs_exact"] = float((np.abs(R - E).max(axis=0) / den).max())
        out["gpu_vs_ref"] = float((np.abs(G - R).max(axis=0) / den).max())
    print(json.dumps(out), flush=True)
