import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import problems as P
from sclmd_b200.engine import MDEngine
ntraj = 1024
nph, nc, ml, nmd, dt = 3000, 300, 4096, 64, 0.25 / 0.658
rng = np.random.default_rng(0)
A = rng.standard_normal((nph, 64)); K = (A @ A.T) / 64 * 0.01
eng = MDEngine(nph, ntraj, dt, nmd); eng.set_dyn(K)
for b in range(2):
    eng.add_bath(list(range(b * 2700, b * 2700 + nc)), P.diag_kernel(ml, nc, dt, b))
    eng.set_noise(b, 0.01 * rng.standard_normal((1, nmd, nc)).repeat(ntraj, 0))
eng.set_state(0.01 * rng.standard_normal((ntraj, nph)), 0.01 * rng.standard_normal((ntraj, nph)), 0)
eng.run(16); eng.set_profiling(True); ms = eng.run(64) / 64
pr = eng.profile_all()
print(json.dumps(dict(variant=os.environ.get("SCLMD_FAR_VARIANT", "0"), ms_per_step=ms, far_ms=pr["tail_far"]["ms"] / max(1, pr["tail_far"]["launches"]),
                      gemm_ms=pr["potforce"]["ms"] / max(1, pr["potforce"]["launches"]))), flush=True)
