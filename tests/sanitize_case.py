"""Small end-to-end case that runs every kernel of the library once.  Two uses:
  * compute-sanitizer (memcheck / racecheck / initcheck, one tool per run), where the GPU pool allows it:
        compute-sanitizer --tool memcheck python tests/sanitize_case.py
  * the SPH_DEBUG_BOUNDS build of the library (asserted kernel indices + guarded allocations, csrc/ctx.cuh):
        SPHB200_LIB=planetmodel-sph_b200/sphb200/libsphb200_dbg.so python tests/sanitize_case.py
    (tests/test_debug_bounds.py does exactly this and requires "guards clean" in the output).
6 000-particle two-body configuration with 64x density contrast (S > 1 stencil, multi-pass cells): tree and direct gravity,
neighbor download, interaction records, diagnostics, field statistics, snapshot; a row-overflow step (max_neighbors = 32);
both optional flags; and a 3-rank group (in-process transport: migration, halo exchange, distributed LBVH, locally essential tree)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))
import sphb200  # noqa: E402
from sphb200 import group as sg  # noqa: E402
from sphb200 import ic  # noqa: E402

c = ic.make_collision(3000, seed=2, separation=2.2, v0=0.3)
n = len(c["h"])
sim = sphb200.Simulation(n, max_neighbors=512)
sim.upload(c["pos"], c["vel"], c["mass"], c["h"])
for impl in (sphb200.GRAVITY_TREE, sphb200.GRAVITY_PARTICLE, sphb200.GRAVITY_TREE):
    sim.step(0.01, impl)
out = sim.download_all()
tree = sim.download_tree()
d = sim.diagnostics()
st = sim.field_stats()
sim.build_neighbors()                     # lists at the current positions: the interaction records need them
off, nbr = sim.download_neighbors()
rec = sim.download_interactions(off, nbr)
with tempfile.TemporaryDirectory() as tmp:
    sim.save_snapshot(os.path.join(tmp, "s.sphb"))
    sim.load_snapshot(os.path.join(tmp, "s.sphb"))
sim.sync()
sim.close()

# rows overflow (32 slots for ~50 neighbors): the truncated-row paths and the stencil re-scan of the density kernel
s2 = sphb200.Simulation(n, max_neighbors=32)
s2.upload(c["pos"], c["vel"], c["mass"], c["h"])
try:
    s2.step(0.01, sphb200.GRAVITY_PARTICLE)
    s2.sync()
except sphb200.SphError as e:
    assert e.code == sphb200.SPH_ERR_NEIGHBOR_OVERFLOW, e
s2.close()

# optional flags: kick-drift, Price & Monaghan softening (direct and tree)
s3 = sphb200.Simulation(n, flags=sphb200.FLAG_KICK_DRIFT | sphb200.FLAG_PM07_SOFTENING | sphb200.FLAG_FIX_KERNEL_DERIV_SIGN)
s3.upload(c["pos"], c["vel"], c["mass"], c["h"])
s3.step(0.01, sphb200.GRAVITY_PARTICLE)
s3.step(0.01, sphb200.GRAVITY_TREE)
s3.sync()
s3.close()

# the decomposition: 3 ranks on one device, several steps (migration), tree and all-pairs gravity
g = sg.Group.single_process(n, [0, 0, 0])
g.upload_global(c["pos"], c["vel"], c["mass"], c["h"])
for impl in (sphb200.GRAVITY_TREE, sphb200.GRAVITY_TREE, sphb200.GRAVITY_PARTICLE, sphb200.GRAVITY_TREE):
    g.step(0.01, impl)
g.sync()
gout = g.download_all()
gd = g.diagnostics()
bad, nalloc = sphb200.debug_check_guards()     # while every allocation is still alive
g.close()
print("sanitize case ok: n=%d neighbors %.1f E_pot %.4f group E_pot %.4f" % (n, len(nbr) / n, d["e_pot"], gd["e_pot"]))
if nalloc < 0:
    print("release build: no guard zones")
else:
    print("guards %s: %d bad bytes over %d allocations" % ("clean" if bad == 0 else "CORRUPTED", bad, nalloc))
    sys.exit(0 if bad == 0 else 3)
