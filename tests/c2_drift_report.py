"""Config C2 (BASELINE.json: 10k-particle sphere, variable smoothing lengths, tree gravity, 100 steps): GPU vs CPU oracle.

Trajectories diverge chaotically once any field differs by an ulp, so the comparison is on aggregates and their drift
(parity protocol P2): total momentum (not conserved by the reference, quirk Q5 -- the criterion is
"same drift as the oracle"), kinetic / potential / internal energy, mean h, mean neighbor count, radial density profile.

    python tests/c2_drift_report.py [steps] > profiles/r01_c2_drift.txt      (needs a B200)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))

import sphb200  # noqa: E402
from sphb200 import ic  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def aggregates(pos, vel, m, rho, grav, h, count, K=1000.0):
    m = m.astype(np.float64)
    return dict(p=(m[:, None] * vel).sum(0), ekin=0.5 * (m * (vel.astype(np.float64) ** 2).sum(1)).sum(),
                epot=0.5 * (m * grav[:, 3]).sum(), eint=(m * K * rho).sum(), hmean=float(h.mean()), nmean=float(count.mean()))


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    dt = 1.0 / 60.0
    c = ic.make_config("c2")
    n = len(c["h"])
    sim = sphb200.Simulation(n)
    sim.upload(c["pos"], c["vel"], c["mass"], c["h"])
    p = sim.effective_params()
    ref = orc.State(c["pos"], c["vel"], c["mass"], c["h"])
    print("# C2: n=%d steps=%d dt=%.5f tree gravity theta=%.2f leaf_max=%d  (GPU libsphb200 vs CPU oracle)" % (n, steps, dt, p.theta, p.leaf_max))
    print("# step | |p| gpu  oracle | E_kin gpu  oracle | E_pot gpu  oracle | E_int gpu  oracle | <h> gpu  oracle | <n> gpu  oracle | median |dx|/R")
    for k in range(1, steps + 1):
        sim.step(dt, sphb200.GRAVITY_TREE)
        orc.step(ref, dt, gravity="tree", max_bits=p.max_grid_bits, leaf_max=p.leaf_max, aabb_mode=p.aabb_mode)
        if k in (1, 2, 5, 10, 20, 50) or k % 25 == 0 or k == steps:
            g = sim.download_all()
            a = aggregates(g["pos"], g["vel"], g["mass"], g["rho"], g["grav"], g["h"], g["count"])
            b = aggregates(ref.pos, ref.vel, ref.mass, ref.rho, ref.grav, ref.h, np.diff(ref.offsets))
            dx = np.median(np.linalg.norm(g["pos"] - ref.pos, axis=1)) / 50.0
            print("%5d | %.6e %.6e | %.6e %.6e | %.6e %.6e | %.6e %.6e | %.5f %.5f | %.3f %.3f | %.2e" % (
                k, np.linalg.norm(a["p"]), np.linalg.norm(b["p"]), a["ekin"], b["ekin"], a["epot"], b["epot"], a["eint"], b["eint"],
                a["hmean"], b["hmean"], a["nmean"], b["nmean"], dx))
    g = sim.download_all()
    r_g = np.linalg.norm(g["pos"], axis=1); r_o = np.linalg.norm(ref.pos, axis=1)
    edges = np.linspace(0, 55, 12)
    print("# radial density profile <rho>(r): bin centre | gpu | oracle")
    for lo, hi in zip(edges[:-1], edges[1:]):
        sg = (r_g >= lo) & (r_g < hi); so = (r_o >= lo) & (r_o < hi)
        print("%6.1f | %.6e | %.6e" % (0.5 * (lo + hi), g["rho"][sg].mean() if sg.any() else 0, ref.rho[so].mean() if so.any() else 0))
    ea = aggregates(g["pos"], g["vel"], g["mass"], g["rho"], g["grav"], g["h"], g["count"])
    eb = aggregates(ref.pos, ref.vel, ref.mass, ref.rho, ref.grav, ref.h, np.diff(ref.offsets))
    rel = lambda x, y: abs(x - y) / max(abs(y), 1e-30)
    print("# final relative differences gpu vs oracle: |p| %.2e  E_kin %.2e  E_pot %.2e  E_int %.2e  <h> %.2e  <n> %.2e" % (
        rel(np.linalg.norm(ea["p"]), np.linalg.norm(eb["p"])), rel(ea["ekin"], eb["ekin"]), rel(ea["epot"], eb["epot"]),
        rel(ea["eint"], eb["eint"]), rel(ea["hmean"], eb["hmean"]), rel(ea["nmean"], eb["nmean"])))


if __name__ == "__main__":
    main()
