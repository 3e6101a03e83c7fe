"""Latency of the reference-facing API on small adaptive problems (the reference's own test sizes: its test script
hints <= 0.12 s for a 2-branch Chebyshev acim, <= 0.3 s for the Fourier ones, test/runtests.jl:18-48)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases
P.acim(cases.lanford())   # context, cuSOLVER handles
out = {}
for name in ("readme_map", "lanford", "runtests_markov_fwd", "runtests_circle_fwd", "circle4", "branches32", "golden_mean_induced"):
    m = getattr(cases, name)()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        L = P.Transfer(m); K = P.SolutionInv(L); rho = K.acim()
        ts.append((time.perf_counter() - t0) * 1e3)
    out[name] = {"acim_ms_min": min(ts), "acim_ms_median": sorted(ts)[2], "coeffs": len(rho.coefficients), "truncation": K.n, "columns_filled": L.datasize}
print(json.dumps(out))
