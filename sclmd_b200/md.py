"""Generalized-Langevin molecular dynamics with the reference's class `md`
(sclmd/md.py:17-794): same constructor, setters, `Run()`, `vv()`, output files.

The time loop runs on the device for `ntraj` independent noise realisations at once
(optional constructor argument, default 1 = the reference's behaviour); trajectories are
reported as the reference's "runs" so `tools.calHF` / `tools.calTC` work unchanged.
Host code here only does set-up (md.setDyn, md.initialise) and file output."""
import os
import sys
import time

import numpy as np

from . import _lib
from numpy import linalg as LA

from . import parallel as PAR
from . import units as U
from .engine import MDEngine
from .functions import bose, chkShape, mdot, symmetrize


class md:
    def __init__(self, dt, nmd, T, syslist=None, axyz=None, dyn=None, nstart=0, nstop=1, npie=1, md2ang=0.06466,
                 ntraj=1, device=None):
        self.nstart, self.nstop = nstart, nstop
        self.dt, self.nmd = dt, nmd
        self.T = T
        self.npie = npie
        # `ntraj` is the size of the WHOLE ensemble.  Inside a torch.distributed job (one process per GPU, torchrun) every rank
        # takes a contiguous block of it: independent trajectories, no data-path collective.  The noise
        # streams are indexed by the global trajectory number, so results do not depend on the number of ranks.
        self.ntraj_global = int(ntraj)
        self.rank, self.world = PAR.rank_world()
        if self.world > 1 and self.ntraj_global >= self.world:
            lo, hi = PAR.shard_range(self.ntraj_global, self.rank, self.world)
            self.traj0, self.ntraj = lo, hi - lo
        else:                               # a single trajectory does not shard: replicas only
            self.traj0, self.ntraj = 0, self.ntraj_global
        self.sharded = self.ntraj != self.ntraj_global
        self.device = PAR.local_device(0) if device is None else int(device)
        self.saveall = self.savep = self.saveq = self.rmnc = False
        self.nstep = None
        self.pforce = None
        self.constraint = None
        self.atomlist = None
        self.SetXyz(axyz)
        if syslist is not None:
            if len(syslist) > self.nta or min(syslist) < 0 or max(syslist) > self.nta - 1:
                print("syslist out of range")
                sys.exit(0)
            self.syslist = np.array(syslist, dtype='int')
            self.na = len(syslist)
            self.nph = 3 * len(syslist)
        elif axyz is not None:
            self.syslist = np.array(range(len(axyz)), dtype='int')
            self.na = len(self.syslist)
            self.nph = 3 * len(self.syslist)
        else:
            self.syslist = self.na = self.nph = None
        self.ml = 1
        self.cf = 0
        self.t = 0
        self.p = []
        self.q = []
        self.pinit = []
        self.qinit = []
        self.baths = []
        self.fhis = []
        self.fbaths = []
        self.etot = np.zeros(nmd)
        self.initranvel = True
        self._eng = None
        self._eng_sig = None
        self._noise_seen = {}
        self.setDyn(dyn)
        self.md2ang = md2ang
        self.mass = []
        if self.els is not None:
            self.get_atommass()
            if len(self.mass) != len(self.els):
                print("Wrong setting in els or mass")
                sys.exit(0)
            self.conv = self.md2ang * np.array([3 * [1.0 / np.sqrt(mass)] for mass in self.mass]).flatten()

    # ------------------------------------------------------------ set-up (host, as in the reference)
    def get_atommass(self):
        for atomsname in self.els:
            if atomsname in U.AtomicMassTable:
                self.mass.append(U.AtomicMassTable[atomsname])

    def info(self):
        print("--------------------------------------------")
        print("Basis information of the MD simulation:")
        print("System atom number:" + str(self.na))
        print("MD time step:" + str(self.dt))
        print("MD number of steps:" + str(self.nmd))
        print("MD memory kernel length:" + str(self.ml))
        print("Number of baths attached:" + str(len(self.baths)))
        print("Trajectories advanced together on the device:" + str(self.ntraj))

    def ResetSavepq(self):
        if self.savep and self.nmd is not None and self.nph is not None:
            self.ps = np.zeros((self.nmd, self.nph))
        if self.saveq and self.nmd is not None and self.nph is not None:
            self.qs = np.zeros((self.nmd, self.nph))

    def energy(self):
        """kinetic energy (md.py:161-165)"""
        return 0.5 * np.sum(np.asarray(self.p) ** 2, axis=-1)

    def AddBath(self, bath):
        """md.py:167-183"""
        if self.dt != bath.dt:
            print("md.AddBath: md time step dt not consistent")
            sys.exit()
        if self.nmd != bath.nmd:
            print("md.AddBath: number of md steps nmd not consistent")
            sys.exit()
        self.baths.append(bath)
        if bath.ml > self.ml:
            self.ml = bath.ml
        self.fbaths.append(np.zeros(self.nph))
        bath._md = self
        bath._index = len(self.baths) - 1
        bath.device = self.device
        self._eng_sig = None

    def AddPowerSection(self, atomlist):
        self.atomlist = atomlist
        self.poweratomlist = np.zeros((len(self.atomlist), self.nmd, 2))

    def AddConstr(self, constr):
        self.constraint = constr
        self._eng_sig = None

    def CalPowerSpec(self, cal=True):
        self.savep = cal
        self.power = np.zeros((self.nmd, 2))       # (the reference leaves it uninitialised; it is only ever weighted with 0 before the first run)

    def CalAveStruct(self, cal=True):
        self.saveq = cal

    def SaveAll(self, save=True):
        self.saveall = save

    def Savep(self, save=True):
        self.savep = save

    def Saveq(self, save=True):
        self.saveq = save

    def SaveTraj(self, nstep=100):
        self.nstep = nstep

    def RemoveNC(self, rmnc=True):
        self.rmnc = rmnc

    def SetT(self, T):
        self.T = T

    def SetMD(self, dt, nmd):
        self.dt, self.nmd = dt, nmd
        self.etot = np.zeros(nmd)
        self._eng_sig = None

    def noranvel(self, rf=False):
        self.initranvel = rf

    def SetXyz(self, axyz):
        if axyz is not None:
            self.xyz = np.array([a[1:] for a in axyz], dtype='d').flatten()
            self.els = [a[0] for a in axyz]
            self.nta = len(axyz)
        else:
            self.xyz = self.els = self.nta = None

    def SetSyslist(self, syslist):
        self.syslist = np.array(syslist)
        self.na = len(syslist)
        self.nph = 3 * len(syslist)
        if self.xyz is not None and len(self.syslist) > self.nta:
            print("md.SetSyslist:system atom number larger than total atom number")
            sys.exit()

    def setDyn(self, dyn=None):
        """md.py:250-292: symmetrise, clip negative eigenvalues, REBUILD dyn = U diag(av) U^T."""
        if dyn is None:
            self.dyn = None
            self.hw = [1.0]
            self.U = None
            self._av = None
            return
        ndyn = np.array(dyn)
        n = chkShape(ndyn)
        if self.nph is not None and self.nph != n:
            print("md.setDyn: the dimension of dynamical matrix is wrong")
            sys.exit(0)
        self.nph = n
        self.dyn = symmetrize(ndyn)
        av, au = LA.eigh(self.dyn)
        if min(av) < 0:
            print("md.setDyn: " + str(int(np.sum(av < 0))) + " negative frequencies removed")
            av = np.where(av < 0, 0.0, av)
        self.hw = np.sqrt(av)
        self.U = np.array(au)
        self._av = np.array(av)            # clipped eigenvalues: the device propagates in this eigenbasis when it can
        self.dyn = mdot(self.U, np.diag(np.array(av)), np.transpose(self.U))
        self._eng_sig = None

    def initialise(self):
        """md.py:294-338: random-phase normal-mode displacements/velocities, one phase set per
        trajectory (trajectory 0 consumes np.random exactly like the reference)."""
        self.t = 0
        shape = (self.nph,) if self.ntraj == 1 else (self.ntraj, self.nph)
        if self.dyn is None or not self.initranvel:
            if self.dyn is not None:
                np.random.rand(len(self.hw))       # the reference draws the phases even when it discards them
            self.p, self.q = np.zeros(shape), np.zeros(shape)
        else:
            av, au = np.asarray(self.hw), self.U
            am = np.zeros(len(av))
            ok = av >= 0.01                          # md.py:317: no motion in slow modes
            am[ok] = np.array([((bose(a, self.T) + 0.5) * 2.0 / a) ** 0.5 for a in av[ok]])
            r = np.random.rand(self.ntraj_global, len(av))
            if self.sharded:                         # rank 0's draw for the whole ensemble, every rank keeps its block
                r = PAR.broadcast_array(r)[self.traj0:self.traj0 + self.ntraj]
            dis = (am * np.cos(2. * np.pi * r)) @ au.T
            vel = -(av * am * np.sin(2. * np.pi * r)) @ au.T
            dis, vel = ApplyConstraint(dis, self.constraint), ApplyConstraint(vel, self.constraint)
            self.p, self.q = vel.reshape(shape), dis.reshape(shape)
        self.pinit, self.qinit = self.p, self.q
        self._state_dirty = True

    def ResetHis(self):
        """md.py:340-349"""
        if self.nph is None or self.ml is None:
            print("self.nph and self.ml are not set")
            sys.exit()
        self._ensure_engine()
        self._eng.reset_history()

    # ------------------------------------------------------------ device plumbing
    def _signature(self):
        # everything that is baked into the device handle: kernels, and for electron baths the bias and the matrices behind the
        # q- and p-dependent forces (ebath.setbias / CheckEmat after AddBath must rebuild it: the reference reads them live, baths.py:243-249)
        def bath_sig(b):
            return (id(b.kernel), b.ml, getattr(b, "bias", None), id(getattr(b, "exim", None)), id(getattr(b, "zeta1", None)),
                    id(getattr(b, "zeta2", None)), id(getattr(b, "efric", None)), tuple(np.asarray(b.cids).tolist()))
        return (self.nph, self.ntraj, self.dt, self.nmd, len(self.baths), id(self.dyn), id(self.pforce),
                tuple(bath_sig(b) for b in self.baths), id(self.constraint))

    def _ensure_engine(self):
        if self.pforce is None and self.dyn is None:
            print("no driver, no md")
            sys.exit()
        sig = self._signature()
        if self._eng is not None and sig == self._eng_sig:
            return
        if self._eng is not None:
            self._eng.close()
        eng = MDEngine(self.nph, self.ntraj, self.dt, self.nmd, self.device)
        if self.pforce is not None:
            # md.py:457-459: a force driver takes precedence over the dynamical matrix.  Its force is a host callback, one call per
            # step (two with constraints); bath forces, history tails, integrator and observables stay on the device.
            eng.set_external_force(True)
        else:
            eng.set_dyn(self.dyn)
            if getattr(self, "_av", None) is not None and self.U is not None and np.shape(self.U) == (self.nph, self.nph):
                try:
                    eng.set_modes(self._av, self.U)
                except _lib.SclmdError:      # md.dyn was reassigned after setDyn: U, hw no longer belong to it -> real space
                    pass
        if self.constraint is not None:
            eng.set_constraint(np.concatenate([np.asarray(list(c), dtype=np.int32) for c in self.constraint]))
        for b in self.baths:
            if b.kernel is None:
                raise RuntimeError("bath kernel is not set: call phbath.gmem() before running (the reference's Run() "
                                   "never does, md.py:493)")
            mq, mp = b._engine_extra()
            eng.add_bath(b.cids, b._engine_kernel(), mq, mp)
        self._eng, self._eng_sig = eng, sig
        self._noise_seen = {}
        self._state_dirty = True
        self._force_out = False

    def _push(self):
        """host attributes -> device (state if it was reassigned, injected noise if it changed)"""
        # reassigned arrays are seen by identity, in-place edits (md.p[:] = ..., md.q *= 0) by a fingerprint of the values
        changed = (getattr(self, "_state_dirty", True) or getattr(self, "_last_p", None) is not self.p or getattr(self, "_last_q", None) is not self.q
                   or getattr(self, "_last_print", None) != self._fingerprint())
        if changed:
            self._eng.set_state(np.asarray(self.q, dtype=float).reshape(self.ntraj, self.nph),
                                np.asarray(self.p, dtype=float).reshape(self.ntraj, self.nph), int(self.t))
            self._state_dirty = False
            self._last_p, self._last_q, self._last_print = self.p, self.q, self._fingerprint()
        for i, b in enumerate(self.baths):
            if b._noise is not None and self._noise_seen.get(i) != b._noise_version:
                nz = np.asarray(b._noise, dtype=float)
                if nz.ndim == 2:
                    nz = np.broadcast_to(nz, (self.ntraj,) + nz.shape)
                elif self.sharded and nz.shape[0] == self.ntraj_global:      # injected for the whole ensemble: keep this rank's block
                    nz = nz[self.traj0:self.traj0 + self.ntraj]
                self._eng.set_noise(i, nz)
                self._noise_seen[i] = b._noise_version

    def _fingerprint(self):
        p, q = np.asarray(self.p, dtype=float), np.asarray(self.q, dtype=float)
        return (p.shape, float(p.sum()), float(np.vdot(p, p)), float(q.sum()), float(np.vdot(q, q)), int(self.t))

    def _pull(self):
        q, p, t = self._eng.get_state()
        shape = (self.nph,) if self.ntraj == 1 else (self.ntraj, self.nph)
        self.q, self.p, self.t = q.reshape(shape), p.reshape(shape), t
        self._last_p, self._last_q, self._last_print = self.p, self.q, self._fingerprint()

    def _device_noise(self, bath):
        """called by bath.gnoi(): fill the device table for every trajectory, no host round trip"""
        self._ensure_engine()
        bath._generate_device_noise(self._eng, bath._index, self.traj0)
        bath._noise_version += 1
        self._noise_seen[bath._index] = bath._noise_version
        if self.ntraj == 1:
            bath._noise = self._eng.get_noise(bath._index)[0]
        else:
            bath._noise = None
        return True

    def get_noise(self, bath_index, traj0=0, ntraj=None):
        """device noise table of one bath: [ntraj, nmd, nc]"""
        self._ensure_engine()
        return self._eng.get_noise(bath_index, traj0, ntraj)

    def _collect(self):
        """bath.cur / md.etot from the device (md.py:383,397)"""
        et = self._eng.etot()
        self.etot = et[0] if self.ntraj == 1 else et
        for i, b in enumerate(self.baths):
            c = self._eng.current(i)
            b.cur = c[0] if self.ntraj == 1 else c

    @property
    def phis(self):
        """md.py:346: [ml, nph] history of p (row 0 newest), rebuilt from the per-bath device rings;
        entries outside the bath dofs are never read by any bath and are reported as 0."""
        self._ensure_engine()
        out = np.zeros((self.ntraj, self.ml, self.nph))
        for i, b in enumerate(self.baths):
            h = self._eng.get_history(i)
            out[:, :b.ml, b.cids] = h
        return out[0] if self.ntraj == 1 else out

    # ------------------------------------------------------------ time stepping
    def _driver_force(self, q):
        """the force driver on every trajectory: driver.force(q[nph]) -> f[nph] (lammpsdriver.py:83-84)"""
        q = np.asarray(q, dtype=float).reshape(self.ntraj, self.nph)
        return np.stack([np.asarray(self.pforce.force(q[k]), dtype=float) for k in range(self.ntraj)])

    def _set_force_output(self, on):
        if self._force_out != on:
            self._eng.set_force_output(on)
            self._force_out = on

    def _advance(self, n):
        if self.pforce is None:
            return self._eng.run(n)
        import time as _time
        t0 = _time.perf_counter()
        for _ in range(n):
            self._eng.step_with_driver(self._driver_force)
        return (_time.perf_counter() - t0) * 1e3

    def vv(self, id=0):
        """one velocity-Verlet step of every trajectory (md.py:367-411), on the device; also fills md.f (force of the last
        evaluation, md.py:411), md.fbaths (bath forces of that evaluation) and, with SaveAll, md.fhis (bath forces of evaluation A, the
        ones the heat current is built from, md.py:397-398)"""
        self._ensure_engine()
        self._push()
        t = int(self.t)
        if self.savep:
            self.ps[t % self.nmd] = np.asarray(self.p).reshape(self.ntraj, self.nph)[0]
        if self.saveq:
            self.qs[t % self.nmd] = np.asarray(self.q).reshape(self.ntraj, self.nph)[0]
        if self.cf:                                       # md.py:378-379
            q0 = np.asarray(self.q, dtype=float).reshape(self.ntraj, self.nph)[0]
            self.cflist.append(np.asarray(self.forcedriver.force(q0)) - self.potforce_harmonic(q0))
        self._set_force_output(True)
        self._advance(1)
        self._pull()
        self._collect()
        f = self._eng.last_force()
        self.f = f[0] if self.ntraj == 1 else f
        for i, b in enumerate(self.baths):
            fb = np.zeros((self.ntraj, self.nph))
            fb[:, b.cids] = self._eng.last_bath_force(i, 2)   # md.fbaths is overwritten by every evaluation: C is what remains
            self.fbaths[i] = fb[0] if self.ntraj == 1 else fb
            if self.saveall:                              # md.py:398 keeps them always; here only when they are dumped
                while len(self.fhis) <= i:
                    self.fhis.append(np.zeros((self.nmd, self.nph)))
                self.fhis[i][t % self.nmd, b.cids] = self._eng.last_bath_force(i, 0)[0]

    def steps(self, n):
        """n steps without host round trips in between (the bulk path used by Run)"""
        self._ensure_engine()
        self._push()
        self._set_force_output(False)
        ms = self._advance(n)
        self._pull()
        return ms

    def CompareForce(self, forcedriver):
        """md.py:362-365: every vv() records forcedriver.force(q) + dyn.q; Run() saves deltaforce.run<j>.npy"""
        self.cf = 1
        self.forcedriver = forcedriver
        self.cflist = []

    def potforce_harmonic(self, q):
        """-dyn . q on the device, whatever driver is attached"""
        q = np.asarray(q, dtype=float)
        f = -_lib.dgemm_nt(q.reshape(-1, self.nph), np.asarray(self.dyn, dtype=float), 1.0, self.device)
        return f[0] if q.ndim == 1 else f

    def force(self, t, p, q, id=0):
        """md.py:413-435 as a stand-alone evaluation for one trajectory (vv()/Run() never call it: the three force evaluations of a
        step are fused into the device kernels).  Histories come from the device ring; products run on the device."""
        from .functions import rpadleft
        it = t + id
        pf = self.potforce(q)
        phis = np.asarray(self.phis)
        phis = phis if phis.ndim == 2 else phis[0]
        qhis = np.zeros_like(phis)
        qhis[0] = np.asarray(q, dtype=float)            # only row 0 of the q history is ever read (baths.py:246-249)
        if id != 0:                                     # md.py:426-431: id = 0 uses the stored history as it is
            phis = rpadleft(phis, p)
        for i in range(len(self.baths)):
            self.fbaths[i] = self.baths[i].bforce(it, phis, qhis)
            pf = pf + self.fbaths[i]
        return pf

    def potforce(self, q):
        """md.py:437-474, harmonic branch: f = -dyn . q on the device (one trajectory or [ntraj, nph])"""
        if self.pforce is not None:
            return self.pforce.force(q)
        if self.dyn is None:
            print("no driver, no md")
            sys.exit()
        return self.potforce_harmonic(q)

    def AddPotential(self, pint):
        """md.py:481-485: a force driver (`.force(q) -> f[nph]`, mass-weighted; lammpsdriver.py:83-84).  The driver runs on the
        host, once per step; everything else of the step stays on the device (sclmd_md_step_begin / sclmd_md_step_end)."""
        self.pforce = pint
        self._eng_sig = None

    def Run(self):
        """md.py:493-682: nstop-nstart runs of nmd steps in npie pieces; noise regenerated per run;
        kappa.* files per run.  With ntraj > 1 trajectory k of run j is reported as run j*ntraj+k."""
        self.initialise()
        self.ResetHis()
        self.info()
        for j in range(self.nstart, self.nstop):
            print("\n" + "MD run: " + str(j))
            tag = (".rank%d" % self.rank) if self.sharded else ""      # one checkpoint per rank of a sharded ensemble
            fn, fnm = "MD" + str(j) + tag + ".nc", "MD" + str(j - 1) + tag + ".nc"
            ipie = -1
            if os.path.isfile(fn):
                ck = self._read_checkpoint(fn)
                ipie = int(ck["ipie"][0])
                if ipie + 1 == self.npie:                 # md.py:535-543
                    print("finished run")
                    if self.savep:
                        self._restore_power(ck)
                    self.t = int(ck["t"][0])
                    continue
                if ipie + 1 > self.npie:                  # md.py:544-547
                    print("ipie error")
                    print("ipie=", ipie)
                    sys.exit()
                print("unfinished run")                   # md.py:515-534
                print("reading resume information")
                self._restore(ck, resume=True)
            else:
                print("new run")
                if os.path.isfile(fnm):
                    print("reading history from previous run")
                    self._restore(self._read_checkpoint(fnm), resume=False)
                elif j == 0 or j == self.nstart:
                    print("initialize a new simulation")
                else:
                    print("no previous nc file exists")
                    sys.exit()
                for b in self.baths:
                    b.gnoi()
                self.ResetSavepq()
            per = int(self.nmd / self.npie)
            trajfile = open('trajectories' + "." + str(self.T) + "." + "run" + str(j) + '.ani', 'w')
            slow = self.savep or self.saveq or self.cf or self.saveall or self.nstep is not None
            for i in range(ipie + 1, self.npie):
                if slow:
                    for _ in range(per):
                        self.vv(j)
                        if self.nstep is not None and ((self.t - 1) == 0 or (self.t - 1) % self.nstep == 0):
                            self._write_frame(trajfile)
                else:
                    self.steps(per)
                self._collect()
                self.dump(i, j)
            trajfile.close()
            if self.cf:                                   # md.py:599-602
                np.save("deltaforce" + ".run" + str(j), np.array(self.cflist) / self.forcedriver.conv)
                self.cflist = []
            if self.savep:                                # md.py:604-653
                power = np.copy(self.power)
                if self.atomlist is not None:
                    poweratomlist = [np.copy(self.poweratomlist[layers]) for layers in range(len(self.atomlist))]
                self.GetPower()
                nrun = j - self.nstart
                self.power = (power * nrun + self.power) / float(nrun + 1)
                if self.atomlist is not None:
                    for layers in range(len(self.atomlist)):
                        self.poweratomlist[layers] = (poweratomlist[layers] * nrun + self.poweratomlist[layers]) / float(nrun + 1)
                self._write_power("power." + str(self.T) + "." + "run" + str(j) + ".dat", self.power)
                if self.atomlist is not None:
                    for layers in range(len(self.atomlist)):
                        self._write_power("poweratomlist." + str(layers) + "." + str(self.T) + "." + "run" + str(j) + ".dat",
                                          self.poweratomlist[layers])
                self.dump(self.npie - 1, j)               # md.py:654-655: dump again, to make sure power is all right
            # heat current (md.py:658-664)
            for ii, b in enumerate(self.baths):
                cur = np.asarray(b.cur).reshape(self.ntraj, self.nmd)
                for k in range(self.ntraj):
                    run = j if self.ntraj_global == 1 else j * self.ntraj_global + self.traj0 + k      # globally unique run index
                    with open("kappa." + str(self.T) + "." + "bath" + str(ii) + ".run" + str(run) + ".dat", "w") as fk:
                        fk.write("%i %f    %f \n" % (run, self.T, np.mean(cur[k]) * U.curcof))
            if self.saveq:
                with open("avestructure." + str(self.T) + "." + "run" + str(j) + ".dat", "w") as f:
                    ave = self.conv * (self.qs.mean(axis=0)) + self.xyz
                    f.write(str(len(self.els)) + '\n' + "average structure" + '\n')
                    for ip in range(len(self.els)):
                        f.write(str(self.els[ip]) + '    ' + str(ave[ip * 3]) + '   ' + str(ave[ip * 3 + 1]) + '   ' + str(ave[ip * 3 + 2]) + '\n')
            if self.rmnc and os.path.exists(fnm):
                os.remove(fnm)

    def mean_currents(self):
        """[nbaths] heat current * curcof averaged over time and over EVERY trajectory of the ensemble (what the kappa files hold):
        per-bath sums on the device, then ONE all-reduce of [sums..., count] over the ranks of a sharded ensemble -- the only
        collective of the path.  Returns (means, global sums, global count)."""
        self._ensure_engine()
        sums = np.array([self._eng.current_sums(i).sum() for i in range(len(self.baths))])
        tot = PAR.allreduce_sum(list(sums) + [float(self.ntraj * self.nmd)]) if self.sharded else np.append(sums, self.ntraj * self.nmd)
        return tot[:-1] / tot[-1] * U.curcof, tot[:-1], int(tot[-1])

    def _write_power(self, name, power):
        """md.py:619-653: only up to 1.5 max(hw)"""
        with open(name, "w") as f:
            for ni in range(len(power)):
                if self.hw is not None and not power[ni, 0] < 1.5 * max(self.hw):
                    break
                f.write("%f     %f \n" % (power[ni, 0], power[ni, 1]))

    def _restore_power(self, ck):
        if "power" in ck:
            self.power = ck["power"]
        if self.atomlist is not None and "poweratomlist" in ck:
            self.poweratomlist = ck["poweratomlist"]

    def _write_frame(self, trajfile):
        q = np.asarray(self.q).reshape(self.ntraj, self.nph)[0]
        fr = np.asarray(self.f).reshape(self.ntraj, self.nph)[0]        # md.py:594-595: element, position, force
        trajfile.write(str(len(self.els)) + '\n' + str(self.t - 1) + '\n')
        structure = self.xyz + self.conv * q
        for ip in range(len(self.els)):
            trajfile.write(str(self.els[ip]) + '    ' + str(structure[ip * 3]) + '   ' + str(structure[ip * 3 + 1]) + '   ' +
                           str(structure[ip * 3 + 2]) + '   ' + str(fr[ip * 3]) + '   ' + str(fr[ip * 3 + 1]) + '   ' + str(fr[ip * 3 + 2]) + '\n')

    def GetPower(self):
        """md.py:351-360 (post-processing; 'next' row of the scope table)"""
        from .tools import powerspecp
        self.power = powerspecp(self.ps, self.dt, self.nmd)
        if self.atomlist is not None:
            for layers in range(len(self.atomlist)):
                self.poweratomlist[layers] = powerspecp(self.ps[:, self.atomlist[layers]], self.dt, self.nmd)

    # ------------------------------------------------------------ checkpoint (md.py:684-745)
    # MD<run>.nc in NetCDF classic format (scipy.io.netcdf_file; netCDF4 is not needed and reads it too) with the reference's
    # variable names: energy, p, q, t, ipie, phis, qhis (+ noise<i>, power, ps, qs when saved).  Extras for the device engine:
    # the per-bath history rings `ring<i>` [ntraj, ml_i, nc_i]; an ensemble adds a leading `ntraj` dimension to p, q, phis, energy.
    def dump(self, ipie, id):
        from scipy.io import netcdf_file
        f = netcdf_file("MD" + str(id) + ((".rank%d" % self.rank) if self.sharded else "") + ".nc", 'w')
        f.title = 'Output from sclmd_b200.md'
        f.createDimension('nph', self.nph)
        f.createDimension('one', 1)
        f.createDimension('two', 2)
        f.createDimension('mem', self.ml)
        f.createDimension('nmd', self.nmd)
        f.createDimension('ntraj', self.ntraj)
        lead = () if self.ntraj == 1 else ('ntraj',)

        def put(name, arr, dims):
            v = f.createVariable(name, 'd', dims)
            v[:] = np.asarray(arr, dtype=float).reshape([f.dimensions[d] for d in dims])
        put('energy', self.etot, lead + ('nmd',))
        put('p', self.p, lead + ('nph',))
        put('q', self.q, lead + ('nph',))
        put('t', [float(self.t)], ('one',))
        put('ipie', [float(ipie)], ('one',))
        phis = np.asarray(self.phis)
        put('phis', phis, lead + ('mem', 'nph'))
        qhis = np.zeros_like(phis)                       # only its first row is ever read (baths.py:246-249), from the live q
        put('qhis', qhis, lead + ('mem', 'nph'))
        for i, b in enumerate(self.baths):
            f.createDimension('n' + str(i), b.nc)
            f.createDimension('m' + str(i), b.ml)
            put('ring' + str(i), self._eng.get_history(i), ('ntraj', 'm' + str(i), 'n' + str(i)))
            put('cur' + str(i), self._eng.current(i), ('ntraj', 'nmd'))       # so that a resumed run averages over ALL its steps
            if self.saveall and self.ntraj == 1 and b._noise is not None:
                put('noise' + str(i), np.asarray(b._noise).reshape(self.nmd, b.nc), ('nmd', 'n' + str(i)))
            seed = getattr(b, "_noise_seed", None)
            if seed is not None:                          # device-generated table: the Philox key regenerates it (global trajectory streams)
                v = f.createVariable('noise_seed' + str(i), 'i', ('two',))
                v[:] = np.array([seed >> 31, seed & 0x7fffffff], dtype=np.int32)
            if self.saveall and i < len(self.fhis):       # md.py:713-714
                put('fhis' + str(i), self.fhis[i], ('nmd', 'nph'))
        if self.savep:
            f.createDimension('npw', len(self.power))
            put('power', self.power, ('npw', 'two'))
            if self.atomlist is not None:                 # md.py:724-726
                f.createDimension('atomlist', len(self.atomlist))
                put('poweratomlist', self.poweratomlist, ('atomlist', 'npw', 'two'))
            if self.saveall:
                put('ps', self.ps, ('nmd', 'nph'))
        if self.saveq and self.saveall:
            put('qs', self.qs, ('nmd', 'nph'))
        f.close()

    @staticmethod
    def _read_checkpoint(fn):
        """MD<run>.nc written by dump() (NetCDF classic).  The reference's own checkpoints are NetCDF-4 / HDF5 with zlib-compressed
        variables (md.py:748-756), which neither scipy nor the package's HDF5 walker (contiguous data only) can read."""
        from scipy.io import netcdf_file
        with open(fn, "rb") as fh:
            if fh.read(4) == b"\x89HDF":
                raise RuntimeError(fn + " is a NetCDF-4 (HDF5) file: checkpoints of the reference package cannot be read here "
                                   "(compressed variables); re-save it in NetCDF classic format, e.g. nccopy -k classic")
        with netcdf_file(fn, 'r', mmap=False) as f:
            return {k: np.array(v[:], dtype=(np.int64 if v.typecode() == 'i' else float)) for k, v in f.variables.items()}

    def _restore(self, ck, resume):
        """state and histories from a checkpoint (md.py:552-562); resume = True: an unfinished run (md.py:515-534) -- also the power
        spectra, the saved p / q series and the noise of the run.  The reference can only continue with saveall, savep and saveq all
        set (the noise series is stored with saveall only); a noise table generated on the device is restored from its Philox key
        instead, whatever the flags."""
        self.p, self.q, self.t = ck["p"], ck["q"], int(ck["t"][0])
        self._ensure_engine()
        self._state_dirty = True
        self._push()
        for i, b in enumerate(self.baths):
            key = "ring%d" % i
            if key in ck and ck[key].shape == (self.ntraj, b.ml, b.nc):
                self._eng.set_history(i, ck[key])
            elif "phis" in ck and self.ntraj == 1 and ck["phis"].shape[0] >= b.ml:      # a checkpoint in the reference's layout
                self._eng.set_history(i, np.ascontiguousarray(ck["phis"][:b.ml][:, b.cids])[None])
        if not resume:
            return
        # heat currents and energies recorded before the interruption (the reference keeps them in memory only, so its resumed runs
        # average the later pieces alone; here the kappa of a resumed run equals that of the uninterrupted one)
        for i, b in enumerate(self.baths):
            if "cur%d" % i in ck:
                self._eng.set_current(i, ck["cur%d" % i])
        if "energy" in ck and ck["energy"].size == self.ntraj * self.nmd:
            self._eng.set_etot(ck["energy"])
        if self.savep:
            self._restore_power(ck)
        self.ResetSavepq()
        missing = (self.savep and "ps" not in ck) or (self.saveq and "qs" not in ck)
        for i, b in enumerate(self.baths):
            if "noise%d" % i in ck:
                b.noise = ck["noise%d" % i]
            elif "noise_seed%d" % i in ck:
                hi, lo = (int(x) for x in ck["noise_seed%d" % i])
                b._regenerate_device_noise(self._eng, i, self.traj0, (hi << 31) | lo)
                b._noise_version += 1
                self._noise_seen[i] = b._noise_version
                b._noise = self._eng.get_noise(i)[0] if self.ntraj == 1 else None
            else:
                missing = True
        if missing:
            print("saveall savep & saveq need to be set true to continue")
            sys.exit(0)
        if self.savep:
            self.ps = ck["ps"]
        if self.saveq:
            self.qs = ck["qs"]


def ApplyConstraint(f, constr=None):
    """md.py:782-794"""
    if constr is None:
        return f
    nf = np.array(f) * 1.0
    for c in constr:
        nf[..., np.asarray(list(c), dtype=int)] = 0
    return nf
