"""On-box probe of the bpt transmission sweep (config-3 shape, n = 483): wall and device-only rate."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import problems as P
from sclmd_b200.negf import bpt

RPC = 6.582119569e-4
K = P.spring_chain_dyn(201, seed=14) / RPC ** 2
b = bpt(None, 0.25, 0.1, [list(range(60, 210)), list(range(393, 543))], [list(range(0, 60)), list(range(543, 603))], dynmatfile=K, num=1000)
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 2960
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
om = np.linspace(0, 0.25 / RPC, nw)
flops = (8 / 3) * 483 ** 3 + 8 * 483 ** 2 * 150
for r in range(reps):
    t0 = time.perf_counter(); tm = b.tm_sweep(om); t1 = time.perf_counter()
    print(json.dumps(dict(nw=nw, wall_s=t1 - t0, omega_per_s=nw / (t1 - t0), tflops_alg=flops * nw / (t1 - t0) / 1e12, checksum=float(tm.sum()))), flush=True)
