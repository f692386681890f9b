import sys, time, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, problems as P
from sclmd_b200.negf import bpt
RPC = 6.582119569e-4
natoms=1000
K = P.spring_chain_dyn(natoms, seed=5) / RPC ** 2
b = bpt(None, 0.25, 0.1, [list(range(0,300)), list(range(2700,3000))], [[],[]], dynmatfile=K, num=10)
om = np.linspace(1.0, 0.25/RPC, 148)
b.tm_sweep(om[:8])
t0=time.perf_counter(); tm=b.tm_sweep(om); t1=time.perf_counter()
n=3000; flops=(8/3)*n**3+8*n*n*300
print(json.dumps(dict(n=n, nw=len(om), s=t1-t0, omega_per_s=len(om)/(t1-t0), tflops_alg=flops*len(om)/(t1-t0)/1e12, finite=bool(np.all(np.isfinite(tm))))))
