"""Physics cross-check of the two pipelines on one junction (the pairing of examples/runmd.py and examples/runnegf.py):
thermal conductance from the ensemble-averaged heat current of quantum-thermostat MD vs. the Landauer value from the NEGF
transmission.  For a harmonic junction with quantum baths the two agree within the statistical error of the MD average."""
import json, os, sys, tempfile, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import problems as P
from sclmd_b200.md import md
from sclmd_b200.baths import ebath
from sclmd_b200.negf import bpt
from sclmd_b200 import units as U

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nmd = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T, delta, dt = 300.0, 0.5, 0.25 / 0.658
natoms = 201
# an ORDERED chain (three polarisations, nearest-neighbour springs): transmission of order one per channel inside the band, so the
# current is well above its thermal fluctuations (the disordered ribbon of the parity tests localises: kappa ~ 1e-3 nW/K)
lap = 2.0 * np.eye(natoms) - np.eye(natoms, k=1) - np.eye(natoms, k=-1)
K = np.kron(lap, np.diag([0.010, 0.006, 0.003])) + 1e-6 * np.eye(3 * natoms)      # hbar w_max = 0.2 / 0.155 / 0.11 eV
fixed = [list(range(0, 60)), list(range(543, 603))]
cats = [list(range(60, 210)), list(range(393, 543))]
damp = 100 / 0.658211814201041          # 0.1 ps in MD time units (examples/runmd.py:45)
os.chdir(tempfile.mkdtemp())
t0 = time.perf_counter()
m = md(dt, nmd, T, axyz=[["C", float(i), 0.0, 0.0] for i in range(natoms)], dyn=K, nstart=0, nstop=2, ntraj=ntraj)
for b, Tb in enumerate((T * (1 + delta / 2), T * (1 - delta / 2))):
    m.AddBath(ebath(cats[b], Tb, dt, nmd, wmax=1.0, nw=500, efric=np.identity(150) / damp))
m.AddConstr(fixed)
m.Run()                                  # run 0 equilibrates, run 1 is measured (state carried over, fresh noise)
cur = [np.asarray(b.cur).reshape(ntraj, nmd) for b in m.baths]
j = np.array([c.mean(axis=1) * U.curcof for c in cur])          # nW per trajectory
jm = (j[0] - j[1]) / 2
kappa_md, err = jm.mean() / (T * delta), jm.std(ddof=1) / np.sqrt(ntraj) / (T * delta)
t1 = time.perf_counter()
RPC = 6.582119569e-4
b = bpt(None, 0.25, 0.1, cats, fixed, dynmatfile=np.array(m.dyn) / RPC ** 2, num=4000)
b.gettm()
kappa_negf = b.thermalconductance(T, delta)
print(json.dumps(dict(ntraj=ntraj, nmd=nmd, kappa_md_nW_per_K=kappa_md, stderr=err, kappa_negf_nW_per_K=kappa_negf,
                      deviation_in_sigma=(kappa_md - kappa_negf) / err, rel=(kappa_md - kappa_negf) / kappa_negf,
                      energy_balance=float((j[0] + j[1]).mean() / jm.mean()), md_s=t1 - t0)))
