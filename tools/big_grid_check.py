import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, time
import poltergeist_jl_b200 as P, cases
from oracle import oracle as O
m=cases.lanford(); n=1<<20
t0=time.time(); G,_=P.engine.transfer_fill(m,n,3,5,chop=False); t1=time.time()
R=O.ref_dense(m.to_desc(),n,3,5,n); t2=time.time()
print("n=2^20: gpu", t1-t0, "s; ref", t2-t1, "s; max err / colnorm", (np.abs(G-R).max(axis=0)/np.linalg.norm(R,axis=0)).max())
