"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck, one tool per run):
    compute-sanitizer --tool memcheck python tests/sanitize_case.py
Runs every kernel of the library once on a 6 000-particle two-body configuration (tree and direct gravity, neighbor
download, interaction records, diagnostics)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))
import sphb200  # noqa: E402
from sphb200 import ic  # noqa: E402

c = ic.make_collision(3000, seed=2, separation=2.2, v0=0.3)
n = len(c["h"])
sim = sphb200.Simulation(n, max_neighbors=512)
sim.upload(c["pos"], c["vel"], c["mass"], c["h"])
for impl in (sphb200.GRAVITY_TREE, sphb200.GRAVITY_PARTICLE, sphb200.GRAVITY_TREE):
    sim.step(0.01, impl)
off, nbr = sim.download_neighbors()
rec = sim.download_interactions(off, nbr)
out = sim.download_all()
tree = sim.download_tree()
d = sim.diagnostics()
sim.sync()
print("sanitize case ok: n=%d neighbors %.1f E_pot %.4f" % (n, len(nbr) / n, d["e_pot"]))
