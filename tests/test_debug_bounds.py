"""Memory-safety evidence without compute-sanitizer (the GPU pool refuses it): the whole hot path -- single handle, row overflow,
optional flags, 3-rank group -- runs on the SPH_DEBUG_BOUNDS build of the library (libsphb200_dbg.so: every data-dependent
shared-memory / row / stack / list index asserted in the kernels, 256-byte guard zones around every device allocation).
Stands in for the reference's job-safety system (its opt-outs: A/Systems/KernelSystem.cs:247, 475, 546)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "planetmodel-sph_b200", "sphb200", "libsphb200_dbg.so")


def test_debug_library_is_built_and_exports_the_abi():
    """CPU: build() produced the debug library and it exports every symbol of the release ABI."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))
    import sphb200
    assert os.path.exists(DBG), "run python __graft_entry__.py (make -C planetmodel-sph_b200/csrc debug)"
    L = ctypes.CDLL(DBG)
    missing = [s for s in sphb200.EXPORTS if not hasattr(L, s)]
    assert not missing, missing
    L.sphb200_version.restype = ctypes.c_char_p
    assert b"SPH_DEBUG_BOUNDS" in L.sphb200_version()


@pytest.mark.gpu
def test_whole_path_on_the_bounds_checked_build():
    env = dict(os.environ, SPHB200_LIB=DBG)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sanitize_case.py")], env=env, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "SPH_DEBUG_BOUNDS" not in out, out[-4000:]          # no kernel assertion fired
    assert "guards clean" in out, out[-4000:]
