import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from sclmd_b200.engine import MDEngine
ntraj, nph = 1024, 3000
rng = np.random.default_rng(0)
A = rng.standard_normal((nph, 64)); K = (A @ A.T) / 64 * 0.01
eng = MDEngine(nph, ntraj, 0.38, 16); eng.set_dyn(K)
eng.set_state(0.01 * rng.standard_normal((ntraj, nph)), 0.01 * rng.standard_normal((ntraj, nph)), 0)
ms = eng.time_potforce(20)
print(json.dumps(dict(variant=os.environ.get("SCLMD_GEMM_VARIANT", "0"), ms=ms, tflops=2.0 * ntraj * nph * nph / ms / 1e9)))
