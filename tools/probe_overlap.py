import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import problems as P
from sclmd_b200.engine import MDEngine
ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nph, nc, ml, nmd, dt = 3000, 300, 4096, 64, 0.25 / 0.658
rng = np.random.default_rng(0)
A = rng.standard_normal((nph, 64)); K = (A @ A.T) / 64 * 0.01
eng = MDEngine(nph, ntraj, dt, nmd); eng.set_dyn(K)
for b in range(2):
    eng.add_bath(list(range(b * 2700, b * 2700 + nc)), P.diag_kernel(ml, nc, dt, b))
    eng.set_noise(b, 0.01 * rng.standard_normal((1, nmd, nc)).repeat(ntraj, 0))
eng.set_state(0.01 * rng.standard_normal((ntraj, nph)), 0.01 * rng.standard_normal((ntraj, nph)), 0)
for blk, ov in ((0, 1), (2, 0), (1, 0), (1, 1)):
    eng.set_tail_block(blk); eng.set_overlap(ov); eng.run(16); eng.set_profiling(True); ms = eng.run(64) / 64
    print(json.dumps(dict(tail_block=blk, overlap=ov, ms_per_step=ms, traj_steps_per_s=ntraj / ms * 1e3, prof=eng.profile_all())), flush=True)
    eng.set_profiling(False)
print(json.dumps(dict(tail_ms=eng.time_tail(0, 10), potforce_ms=eng.time_potforce(10))))
