"""Quick on-box probe: FP64 peaks, DGEMM and tail-kernel throughput at C5-like sizes."""
import ctypes as C
import json
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from sclmd_b200 import _lib
from sclmd_b200.engine import MDEngine

out = {}
L = _lib.lib()
for kind, name in ((0, "dfma_tflops"), (1, "dmma_tflops")):
    v = C.c_double(0)
    _lib.check(L.sclmd_probe_fp64(0, kind, C.byref(v)))
    out[name] = v.value
try:
    import torch
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2):
        (a @ b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["cublas_dgemm_8192_tflops"] = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a, b, c
    torch.cuda.empty_cache()
except Exception as ex:  # noqa
    out["cublas_error"] = repr(ex)
print(json.dumps(out), flush=True)

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nph, nc, ml, nmd, dt = 3000, 300, 4096, 64, 0.25 / 0.658
rng = np.random.default_rng(0)
A = rng.standard_normal((nph, 64))
K = (A @ A.T) / 64 * 0.01
eng = MDEngine(nph, ntraj, dt, nmd)
eng.set_dyn(K)
import problems as P
for b in range(2):
    eng.add_bath(list(range(b * 2700, b * 2700 + nc)), P.diag_kernel(ml, nc, dt, b))
    eng.set_noise(b, 0.01 * rng.standard_normal((1, nmd, nc)).repeat(ntraj, 0))
eng.set_state(0.01 * rng.standard_normal((ntraj, nph)), 0.01 * rng.standard_normal((ntraj, nph)), 0)
eng.run(3)
ms = eng.run(10) / 10
tail_ms = eng.time_tail(0, 5)
pf_ms = eng.time_potforce(5)
bytes_tail = ntraj * (ml - 1) * nc * 8.0
res = dict(ntraj=ntraj, ms_per_step=ms, traj_steps_per_s=ntraj / (ms * 1e-3), tail_ms=tail_ms,
           tail_GBs=bytes_tail / (tail_ms * 1e-3) / 1e9, potforce_ms=pf_ms,
           potforce_tflops=2.0 * ntraj * nph * nph / (pf_ms * 1e-3) / 1e12, launches=eng.launch_count())
print(json.dumps(res), flush=True)
eng.close()
