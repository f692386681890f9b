"""Stage timing of the device noise generator at the config-5 bath shape (nc = 300, nmd = 8192, 1024 trajectories):
normal draws, x = L xi (batched DMMA GEMM), mirrored transform; and the per-frequency factorisation at nc = 150 / 300.
usage: python tools/probe_noise.py [ntraj] [nc] [nmd]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from sclmd_b200 import _lib, noise as N           # noqa: E402
from sclmd_b200.engine import MDEngine            # noqa: E402
import problems as P                              # noqa: E402

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nmd = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
dt = 0.25 / 0.658
out = {"ntraj": ntraj, "nc": nc, "nmd": nmd}
nph = 2 * nc + 6
eng = MDEngine(nph, ntraj, dt, nmd)
eng.set_dyn(np.eye(nph))
eng.add_bath(list(range(nc)), np.full((1, nc), 0.01))
gam = np.array([np.eye(nc) * 0.05 * np.pi / 6.0])
plan = N.ph_plan(gam, np.array([0.0]), 300.0, 0.5, dt, nmd)
for rep in range(2):
    t0 = time.perf_counter()
    _lib.check(_lib.lib().sclmd_md_generate_noise(eng._h, 0, plan._h, C.c_uint64(7), 0))
    wall = time.perf_counter() - t0
    pr = plan.profile()
    pr["wall_s"] = wall
    pr["samples_per_s"] = ntraj * nmd * nc / wall
    pr["frac_of_hbm_roofline_16B_per_sample"] = pr["samples_per_s"] * 16 / 6459e9
    out["generate_rep%d" % rep] = pr
plan.close()
eng.close()
# per-frequency factorisation (a different matrix for every frequency)
for fnc, fnmd in ((150, 4096), (300, 512)):
    gwl, g = P.gamma_grid(3, fnc, 9, wmax=0.9 * np.pi / dt)
    t0 = time.perf_counter()
    pl = N.ph_plan(g, gwl, 300.0, 0.95 * np.pi / dt, dt, fnmd)
    wall = time.perf_counter() - t0
    out["factor_nc%d" % fnc] = {"nfactors": fnmd // 2 + 1, "factor_ms": pl.profile()["factor_ms"], "wall_s_incl_host_setup": wall}
    pl.close()
print(json.dumps(out))
