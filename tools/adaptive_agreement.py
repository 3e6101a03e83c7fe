"""How often the GPU cached operator and the faithful restatement choose the same grid / colstop per column, and why not."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import poltergeist_jl_b200 as P
import cases
from oracle import oracle as O
import adaptive_explain as AE
CASES = [("lanford", 120), ("readme_map", 100), ("circle4", 129), ("runtests_circle_rev", 65), ("runtests_circle_fwd", 65),
         ("runtests_markov_rev", 64), ("runtests_markov_fwd", 64), ("sin_asin_map", 100), ("branches32", 200)]
for name, ncols in CASES:
    m = getattr(cases, name)()
    desc = m.to_desc()
    L = P.Transfer(m); L.resizedata(ncols)
    R = O.RefTransfer(desc).resizedata(ncols)
    cs = np.array([L.colstop(k) for k in range(ncols)]); rs = np.array([R.colstop(k) for k in range(ncols)])
    ng = L.nodes_used(0, ncols); nr = R.nodes_used[:ncols]
    bad = [k for k in range(ncols) if cs[k] != rs[k] or ng[k] != nr[k]]
    print(json.dumps({"case": name, "columns": ncols, "grids_equal": int((ng == nr).sum()), "colstops_equal": int((cs == rs).sum()), "mismatched": len(bad)}))
    for k in bad:
        ok, txt = AE.explain_column(P, O, m, desc, k, int(ng[k]), int(cs[k]), int(nr[k]), int(rs[k]))
        print(f"   col {k}: explained={ok}  {txt}")
