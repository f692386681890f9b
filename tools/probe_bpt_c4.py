"""On-box probe of the config-4 shaped NEGF sweeps (n = 654, biased block of 36 dofs): transmission and biased power spectrum."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import problems as P
from sclmd_b200.negf import bpt

RPC = 6.582119569e-4
natoms = 242
K = P.spring_chain_dyn(natoms, seed=15) / RPC ** 2
fixed = [list(range(0, 24)), list(range(678, 726))]
bath = [list(range(24, 144)), list(range(558, 678))]
b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=10)
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
om = np.linspace(0.5, 0.25 / RPC, nw)
n = len(b.dynmat)
for r in range(2):
    t0 = time.perf_counter(); tm = b.tm_sweep(om); t1 = time.perf_counter()
flops = (8 / 3) * n ** 3 + 8 * n ** 2 * 120
print(json.dumps(dict(what="tm n=654", nw=nw, wall_s=t1 - t0, omega_per_s=nw / (t1 - t0), tflops_alg=flops * nw / (t1 - t0) / 1e12)), flush=True)
lam = P.c4_lambda()
sl = list(range(111 * 3, 123 * 3))
b.setbias(1.0, bdamp=lam["eta_r"] / RPC, chiplus=lam["xip_r"] / RPC ** 2, chiminus=lam["xim_r"] / RPC ** 2, dofatomofbias=sl)
for r in range(2):
    t0 = time.perf_counter(); ps = b.ps_sweep(om, 300.0, sl); t1 = time.perf_counter()
flops = 2 * ((8 / 3) * n ** 3 + 8 * n ** 2 * 36)
print(json.dumps(dict(what="ps (biased, two factorisations) n=654", nw=nw, wall_s=t1 - t0, omega_per_s=nw / (t1 - t0),
                      tflops_alg=flops * nw / (t1 - t0) / 1e12, finite=bool(np.all(np.isfinite(ps))))), flush=True)
