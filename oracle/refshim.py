"""TEST INFRASTRUCTURE ONLY -- imports the read-only reference (ydsbbt/sclmd) in place.

Used by ``oracle/make_golden.py`` (run in the build container, where
``/root/reference`` exists) to validate ``oracle/sclmd_oracle.py`` and to write
the fixtures under ``tests/golden/``.  Nothing in the product package
(``sclmd_b200``), ``bench.py`` or the ``-m gpu`` tests may import this module:
``/root/reference`` does not exist on the GPU box.

The reference needs three things to import under NumPy 2 without LAMMPS /
netCDF4:
  * ``netCDF4.Dataset``                (sclmd/md.py:9, sclmd/myio.py:7)
  * ``lammps.lammps``                  (sclmd/negf.py:5, sclmd/lammpsdriver.py:10)
  * removed aliases ``np.complex`` etc (sclmd/baths.py:205, sclmd/negf.py:154)
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SCLMD_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sclmd"))


def install():
    """Make ``import sclmd`` resolve to the reference tree; returns the module dict."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name, typ in (("complex", complex), ("float", float), ("int", int),
                      ("complex_", np.complex128)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    if "netCDF4" not in sys.modules:
        nc4 = types.ModuleType("netCDF4")

        class Dataset:  # write-only no-op so md.dump() works
            def __init__(self, *a, **k):
                self.variables = {}

            def createDimension(self, *a, **k):
                pass

            def createVariable(self, name, *a, **k):
                class V:
                    def __setitem__(s, key, val):
                        s.v = np.array(val)
                self.variables[name] = V()
                return self.variables[name]

            def close(self):
                pass
        nc4.Dataset = Dataset
        sys.modules["netCDF4"] = nc4
    if "lammps" not in sys.modules:
        lm = types.ModuleType("lammps")

        class lammps:
            def __init__(self, *a, **k):
                raise RuntimeError("LAMMPS is not available")
        lm.lammps = lammps
        sys.modules["lammps"] = lm
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sclmd.md as rmd
    import sclmd.baths as rbaths
    import sclmd.noise as rnoise
    import sclmd.functions as rfun
    import sclmd.negf as rnegf
    import sclmd.selfenergy as rsig
    import sclmd.units as runits
    ident = lambda it, *a, **k: it
    rmd.tqdm = ident
    rnoise.tqdm = ident
    return dict(md=rmd, baths=rbaths, noise=rnoise, functions=rfun,
                negf=rnegf, selfenergy=rsig, units=runits)


@contextlib.contextmanager
def quiet():
    """The reference prints a lot; swallow it."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield


def make_bpt(mods, dynmat, damp, dofatomofbath, dofatomfixed, natoms, maxomega, num):
    """Construct reference ``bpt`` without LAMMPS (negf.py:10-24 needs lammps())."""
    o = object.__new__(mods["negf"].bpt)
    o.rpc = 6.582119569e-4
    o.bc = 8.617333262e-5
    o.damp = damp
    o.maxomega = maxomega / o.rpc
    o.intnum = num
    o.dofatomfixed = dofatomfixed
    o.isbias = False
    o.dofatomofbias = []
    o.dofatomofbath = dofatomofbath
    o.natoms = natoms
    o.dynmat = np.array(dynmat)
    return o


def make_sig(mods, K00, K11, K01, maxomega, num, eta=0.164e-3):
    """Construct reference ``sig`` without LAMMPS (selfenergy.py:9-26)."""
    o = object.__new__(mods["selfenergy"].sig)
    o.rpc = 6.582119569e-4
    o.maxomega = maxomega / o.rpc
    o.intnum = num
    o.eta = eta / o.rpc
    o.ep = np.linspace(0, o.maxomega, o.intnum + 1)
    o.K00 = np.array(K00)
    o.K11 = np.array(K11)
    o.K01 = (np.array(K01) + np.array(K01)) / 2
    o.K10 = np.transpose(o.K01)
    return o
