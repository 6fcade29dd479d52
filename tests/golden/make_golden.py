#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded inputs + the CPU oracle's outputs after one reference step.

The reference (C#/Unity) cannot run here, so these vectors are *oracle* outputs, not reference outputs; they (a) freeze
the oracle against silent drift (tests/test_golden.py re-runs it and demands bit-equality) and (b) give the GPU tests a
fixture that does not depend on the oracle being rebuilt.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))

from sphb200 import ic  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = {
    # name: (n, gravity, dt, h scale, max_bits, leaf_max)
    "sphere256_direct": (256, "direct", 1.0 / 60.0, 2.0, 3, 4),
    "sphere512_tree": (512, "tree", 0.02, 2.0, 3, 4),
}


def build(name):
    n, gravity, dt, hs, max_bits, leaf_max = CASES[name]
    c = ic.make_sphere(n, seed=777)
    c["h"] = (c["h"] * np.float32(hs)).astype(np.float32)
    c["vel"] = np.random.Generator(np.random.Philox(99)).normal(0, 0.4, c["pos"].shape).astype(np.float32)
    s = orc.State(c["pos"], c["vel"], c["mass"], c["h"])
    orc.step(s, dt, gravity=gravity, max_bits=max_bits, leaf_max=leaf_max, accum_double=False)
    own1 = s.n_own.copy()
    out = dict(in_pos=c["pos"], in_vel=c["vel"], in_mass=c["mass"], in_h=c["h"], dt=np.float32(dt),
               max_bits=np.int32(max_bits), leaf_max=np.int32(leaf_max), offsets=s.offsets, nbr=s.nbr, n_own=own1,
               rho=s.rho, P=s.P, gradP=s.gradP, grav=s.grav, pos=s.pos, vel=s.vel, h=s.h,
               num_particles=s.num_particles, num_approx=s.num_approx)
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
