"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/sphb200.h declares, and the component structs are byte-exact mirrors of the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sphb200.h")).read()
    return sorted(set(re.findall(r"SPH_API\s+[\w\s\*]+?\b(sphb200_\w+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    import sphb200
    L = sphb200.load_library()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    assert sorted(sphb200.EXPORTS) == syms                    # python binding list == header
    assert b"sm_100a" in L.sphb200_version()


def test_default_params_are_the_reference_constants():
    import sphb200
    p = sphb200.default_params()
    assert p.K == 1000.0            # PressureFieldSystem.cs:31
    assert p.G == 1.0               # GravityFieldSystem.cs:26
    assert p.theta == np.float32(0.7)   # GravityFieldSystem.cs:228
    assert p.target_neighbors == 50  # ParticleSmoothingSystem.cs:18
    assert p.leaf_max == 4 and p.aabb_mode == 0 and p.flags == 0
    assert C.sizeof(sphb200.Params) == 48 and C.sizeof(sphb200.GridParams) == 40


def test_component_struct_layouts():
    import sphb200
    assert sphb200.Translation.itemsize == 12
    assert sphb200.PhysicsVelocity.itemsize == 24 and sphb200.PhysicsVelocity.fields["angular"][1] == 12
    s = sphb200.ParticleSmoothing
    assert s.itemsize == 28 and s.fields["supportDomain"][1] == 4 and s.fields["sphereColliderPosRadius"][1] == 8
    assert s.fields["neighbors"][1] == 24
    g = sphb200.GravityField
    assert g.itemsize == 24 and g.fields["numParticles"][1] == 16 and g.fields["numApprox"][1] == 20
    i = sphb200.ParticleInteraction
    assert i.itemsize == 40 and i.fields["kernelThis"][1] == 8 and i.fields["kernelSymmetric"][1] == 24


def test_header_struct_sizes_match_python_mirrors(tmp_path):
    """Compile a tiny C program against include/sphb200.h and compare sizeof/offsetof with the numpy mirrors."""
    import subprocess
    import sphb200
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sphb200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(sph_Translation),sizeof(sph_PhysicsVelocity),sizeof(sph_ParticleSmoothing),sizeof(sph_GravityField),"
                   "sizeof(sph_ParticleInteraction),sizeof(sph_Params),sizeof(sph_GridParams),offsetof(sph_ParticleSmoothing,neighbors),"
                   "offsetof(sph_GravityField,numParticles));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).split()
    assert [int(x) for x in out] == [12, 24, 28, 24, 40, 48, 40, 24, 16]


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device create() must fail with SPH_ERR_CUDA -- there is no CPU path in the product."""
    import torch
    import sphb200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sphb200.SphError) as e:
        sphb200.Simulation(128)
    assert e.value.code == sphb200.SPH_ERR_CUDA


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "planetmodel-sph_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".cs")):
                txt = open(os.path.join(dp, f)).read()
                assert "liborc" not in txt and "sph_oracle" not in txt.replace("oracle/sph_oracle.cpp", ""), f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
