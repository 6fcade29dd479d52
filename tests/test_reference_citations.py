"""Every `File.cs:line` / `File.cs:first-last` citation of the reference in the headers, sources, docs and tests points at a file
that exists under /root/reference and at lines inside it.  The reference tree is only present in the build container: the
test skips itself elsewhere (nothing at run time reads /root/reference)."""
import collections
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

CITED_FROM = (["include/sphb200.h", "DESIGN.md", "INTEGRATION.md", "README.md", "oracle/sph_oracle.cpp", "oracle/oracle.py", "bench.py"]
              + ["planetmodel-sph_b200/csrc/*.cu", "planetmodel-sph_b200/csrc/*.cuh", "planetmodel-sph_b200/sphb200/*.py",
                 "planetmodel-sph_b200/csharp/*.cs", "planetmodel-sph_b200/host_cpp/*pp", "tests/*.py"])
CITATION = re.compile(r"([A-Za-z_][\w\.]*\.(?:cs|md|json|unity)):(\d+)(?:-(\d+))?")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists only in the build container")
def test_every_cited_reference_line_exists():
    index = collections.defaultdict(list)
    for dp, _, fs in os.walk(REF):
        for f in fs:
            if f.endswith((".cs", ".md", ".json", ".unity")):
                index[f].append(os.path.join(dp, f))
    lengths = {}
    checked, bad = 0, []
    for patt in CITED_FROM:
        for path in sorted(glob.glob(os.path.join(ROOT, patt))):
            txt = open(path, errors="replace").read()
            for m in CITATION.finditer(txt):
                name = m.group(1).split("/")[-1]
                if name not in index:          # abbreviated names ("...BuilderTests.cs") and this repo's own files
                    continue
                first, last = int(m.group(2)), int(m.group(3) or m.group(2))
                for c in index[name]:
                    if c not in lengths:
                        lengths[c] = sum(1 for _ in open(c, errors="replace"))
                longest = max(lengths[c] for c in index[name])
                checked += 1
                if not (1 <= first <= last <= longest):
                    bad.append((os.path.relpath(path, ROOT), m.group(0), longest))
    assert checked > 250, checked
    assert not bad, bad
    # the six systems and the kernel file are cited from the C header itself (the boundary names what it replaces)
    hdr = open(os.path.join(ROOT, "include", "sphb200.h")).read()
    for f in ("ParticleSmoothingSystem.cs", "KernelSystem.cs", "GravityFieldSystem.cs", "DensityFieldSystem.cs",
              "PressureFieldSystem.cs", "VelocitySystem.cs", "BuildPhysicsWorld.cs", "ExportPhysicsWorld.cs", "Integrator.cs"):
        assert re.search(re.escape(f) + r":\d+", hdr), f
