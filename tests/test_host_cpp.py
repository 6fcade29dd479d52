"""The compiled C++ host mirror (host_cpp/sph_systems.hpp: the six systems + World, and GroupWorld over sphb200_group_*) ticking
the reference scene through the C ABI -- the stand-in for the C# host, which cannot be built here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "planetmodel-sph_b200", "host_cpp", "host_demo")


def _run(args, env=None):
    r = subprocess.run([DEMO] + [str(a) for a in args], capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def _energies(out):
    rows = [tuple(float(x) for x in re.findall(r"mass (\S+)  \|p\| (\S+)  E_kin (\S+)  E_pot (\S+)  E_int (\S+)", ln)[0])
            for ln in out.splitlines() if ln.startswith("step")]
    assert rows
    return rows


@pytest.mark.gpu
def test_cpp_host_ticks_the_six_systems_and_the_group_agrees():
    single = _energies(_run([3000, 4, "tree"]))
    assert len(single) == 4 and abs(single[0][0] - 100.0) < 1e-3            # total mass of the reference scene
    group = _energies(_run([3000, 4, "tree", 3], env={"SPH_DEMO_ONE_GPU": "1"}))   # 3 ranks on one GPU: in-process transport
    assert len(group) == 4
    for a, b in zip(single, group):                                          # tree gravity: the group is bit-identical, sums to fp64 noise
        for x, y in zip(a, b):
            assert abs(x - y) <= 1e-9 * abs(x) + 1e-12
    direct = _energies(_run([3000, 2, "particle"]))
    assert abs(direct[0][3] - single[0][3]) <= 2e-2 * abs(single[0][3])      # tree vs direct potential energy: per cent level
