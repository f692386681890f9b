"""Single-trajectory stepping rate of the config-1 shape (603 dofs, two ml = 1 diagonal baths, fixed ends): persistent kernel vs launch chain."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import bench
w = dict(bench.WORKLOADS["c2"])
for ntraj in (1, 2, 4):
    eng, _ = bench.make_engine(w, 0, ntraj)
    rng = np.random.default_rng(1)
    for b in range(2):
        eng.set_noise(b, 0.01 * rng.standard_normal((ntraj, w["nmd"], w["nc"])))
    nph = 3 * w["natoms"]
    eng.set_state(0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph)), 0)
    eng.run(64)
    ms = eng.run(4096)
    print(json.dumps(dict(ntraj=ntraj, persist=os.environ.get("SCLMD_NO_PERSIST") is None, us_per_step=ms / 4096 * 1e3,
                          steps_per_s=4096 / (ms * 1e-3), launches=eng.launch_count())), flush=True)
    eng.close()
