"""On-box throughput probe for the NEGF / sig sweeps and the noise generator."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import problems as P
from sclmd_b200.negf import bpt
from sclmd_b200.selfenergy import sig
from sclmd_b200 import noise as N

RPC = 6.582119569e-4
out = {}
K = P.spring_chain_dyn(201, seed=14) / RPC ** 2
b = bpt(None, 0.25, 0.1, [list(range(60, 210)), list(range(393, 543))], [list(range(0, 60)), list(range(543, 603))], dynmatfile=K, num=1000)
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
om = np.linspace(0, 0.25 / RPC, nw)
b.tm_sweep(om[:296])
t0 = time.perf_counter(); tm = b.tm_sweep(om); t1 = time.perf_counter()
flops = (8 / 3) * 483 ** 3 + 8 * 483 ** 2 * 150
out["bpt_tm"] = dict(nw=nw, s=t1 - t0, omega_per_s=nw / (t1 - t0), tflops_alg=flops * nw / (t1 - t0) / 1e12)
print(json.dumps(out), flush=True)
m = 24
K00, K11, K01 = P.chain_blocks(m, seed=5, k2=0.08)
full = np.zeros((2 * m, 2 * m)); full[:m, :m], full[m:, m:], full[:m, m:], full[m:, :m] = K00, K11, K01, K01.T
s = sig(None, 0.12, range(0, m), range(m, 2 * m), dynmatfile=full, num=20000, eta=0.164e-3)
s.selfenergy_sweep(s.ep[:100], 'R')
t0 = time.perf_counter(); se = s.selfenergy_sweep(s.ep, 'R'); t1 = time.perf_counter()
out["sig_getse"] = dict(nw=len(s.ep), s=t1 - t0, omega_per_s=len(s.ep) / (t1 - t0), mean_iters=float(s.iterations.mean()))
t0 = time.perf_counter(); tmv = s.tm_sweep(s.ep); t1 = time.perf_counter()
out["sig_gettm"] = dict(nw=len(s.ep), s=t1 - t0, omega_per_s=len(s.ep) / (t1 - t0))
print(json.dumps(out), flush=True)
# noise: C5 bath (Debye-like single basis) and a gamma-grid bath at nc=150
for name, nc, nmd, ngw, ntraj in (("noise_c5_single_basis", 300, 8192, 1, 64), ("noise_grid_nc150", 150, 4096, 8, 64), ("noise_grid_nc300", 300, 512, 4, 8)):
    gwl, gam = P.gamma_grid(max(ngw, 2), nc, 3)
    if ngw == 1:
        gwl, gam = np.array([0.0]), gam[:1]
    t0 = time.perf_counter(); plan = N.ph_plan(gam, gwl, 300.0, 1.0, 0.38, nmd); t1 = time.perf_counter()
    x = plan.generate(2, seed=1)
    t2 = time.perf_counter(); x = plan.generate(ntraj, seed=1); t3 = time.perf_counter()
    out[name] = dict(nc=nc, nmd=nmd, plan_s=t1 - t0, gen_s=t3 - t2, samples_per_s=ntraj * nmd * nc / (t3 - t2))
    plan.close()
    print(json.dumps(out[name]), flush=True)
