import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import problems as P
from oracle import sclmd_oracle as O
from sclmd_b200.engine import MDEngine

def run(kind, ml, nc, ntraj, cons_on=True):
    natoms = max(12, (2 * nc + 8) // 3 + 2); nph = 3 * natoms
    dt, nmd = 0.25 / 0.658, 64
    K = P.psd_project(P.spring_chain_dyn(natoms, seed=5))
    cons = [list(range(0, 3)), list(range(nph - 3, nph))] if cons_on else None
    cids = [list(range(3, 3 + nc)), list(range(nph - 3 - nc, nph - 3))]
    eng = MDEngine(nph, ntraj, dt, nmd); ens = O.EnsembleMD(K, dt, nmd, ntraj, cons)
    eng.set_dyn(K)
    if cons_on: eng.set_constraint([i for g in cons for i in g])
    for b in range(2):
        kern = P.diag_kernel(ml, nc, dt, 50 + b) if kind == "diag" else P.full_kernel(ml, nc, dt, 50 + b)
        nz = P.injected_noise(ntraj, nmd, nc, seed=60 + b)
        eng.add_bath(cids[b], kern); eng.set_noise(b, nz); ens.add_bath(cids[b], kern, nz)
        print("noise roundtrip", np.abs(eng.get_noise(b) - nz).max())
    rng = np.random.default_rng(7)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    eng.set_state(q0, p0, 0); ens.q[:], ens.p[:] = q0, p0
    for s in range(3):
        eng.run(1); ens.run(1)
        q, p, t = eng.get_state()
        eq = np.abs(q - ens.q); ep = np.abs(p - ens.p)
        print(kind, ml, nc, ntraj, "step", s, "q err per traj", eq.max(1), "argmax dof", eq.argmax(1), "p err", ep.max(1), ep.argmax(1))
    eng.close()
run("diag", 300, 30, 7)
run("diag", 2, 6, 1)
run("diag", 2, 6, 1, cons_on=False)
run("full", 2, 6, 1)
