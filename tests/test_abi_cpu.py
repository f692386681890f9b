"""CPU-only checks of the boundary: the C-ABI library is built, loads, and exports every
symbol include/sclmd_b200.h declares.  No compute call is made (no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sclmd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sclmd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sclmd_b200 import build, _lib
    build.build()
    L = _lib.lib()
    decl = header_symbols()
    assert len(decl) >= 20
    missing = [s for s in decl if not hasattr(L, s)]
    assert not missing, "declared in include/sclmd_b200.h but not exported: %s" % missing
    unbound = [s for s in decl if s not in _lib.exported_symbols()]
    assert not unbound, "declared but not bound in sclmd_b200/_lib.py: %s" % unbound


def test_library_is_sm100a_only():
    from sclmd_b200 import build
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    from sclmd_b200 import _lib
    try:
        n = _lib.device_count()
    except _lib.SclmdError:
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is present")
    from sclmd_b200.engine import MDEngine
    with pytest.raises(_lib.SclmdError):
        MDEngine(6, 1, 0.1, 8)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sclmd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "sclmd_oracle" not in txt, f


def test_error_codes_map_to_reference_exceptions():
    """sig.sgf raises ValueError after 100 decimation iterations (selfenergy.py:127-130):
    SCLMD_ERR_NOCONV surfaces as a ValueError subclass; everything else as SclmdError."""
    from sclmd_b200 import _lib
    with pytest.raises(ValueError):
        _lib.check(-4)
    with pytest.raises(_lib.SclmdError):
        _lib.check(-1)
    assert _lib.check(0) == 0
