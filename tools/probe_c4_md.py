"""Throughput of the config-4 shape (examples/current-induced/rundp.py: 726 dofs, 72 fixed, two 120-dof electron baths with scalar
friction and a biased 36-dof bath with the example's friction / non-conservative / Berry matrices) as an ensemble through the
reference-shaped classes:  python tools/probe_c4_md.py [ntraj] [nsteps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P                                   # noqa: E402
from sclmd_b200.md import md                           # noqa: E402
from sclmd_b200.baths import ebath                     # noqa: E402

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 512
c = P.md_case_c4_shape()
nmd = 2048
natoms = c["K"].shape[0] // 3
m = md(c["dt"], nmd, c["T"], axyz=[["C", float(i), 0.0, 0.0] for i in range(natoms)], dyn=c["K"], ntraj=ntraj)
e = c["e"]
rng = np.random.default_rng(1)
for b in range(3):
    bb = ebath(c["cids"][b], c["T"], c["dt"], nmd, wmax=2.0, nw=50, bias=e["bias"][b], efric=e["efric"][b], exim=e["exim"][b],
               exip=e["exip"][b], zeta1=e["zeta1"][b], zeta2=e["zeta2"][b], zpmotion=False)
    bb.noise = 0.003 * rng.standard_normal((nmd, len(c["cids"][b])))
    m.AddBath(bb)
m.AddConstr(c["cons"])
m.noranvel()
m.initialise()
m.ResetHis()
m.steps(64)
t0 = time.perf_counter()
ms = m.steps(nsteps)
wall = time.perf_counter() - t0
print("c4 shape: ntraj %d, %d steps: device %.3f ms/step, %.3e trajectory-steps/s (wall %.3e)"
      % (ntraj, nsteps, ms / nsteps, ntraj * nsteps / (ms * 1e-3), ntraj * nsteps / wall), file=sys.stderr)
import json                                            # noqa: E402
print(json.dumps({"metric": "qtb_md_trajectory_steps_per_s", "value": ntraj * nsteps / (ms * 1e-3), "unit": "trajectory-steps/s",
                  "ms_per_step": ms / nsteps, "wall_value": ntraj * nsteps / wall, "ntraj": ntraj, "steps": nsteps,
                  "config": "BASELINE configs[3] shape (examples/current-induced/rundp.py): 726 dofs, 72 fixed, two 120-dof electron "
                            "baths with scalar friction + the biased 36-dof bath with the example's friction / non-conservative / "
                            "Berry matrices (tests/golden/c4_lambda.npz), ml = 1, through the md / ebath classes"}))
