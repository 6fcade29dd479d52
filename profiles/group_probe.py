"""Profiling probe: a group of W ranks that share cuda:0 (in-process transport) stepping a scaled C4 sphere, so that the
decomposition kernels (kernels_group.cu, kernels_sort.cu) can be captured with ncu on a single-GPU box:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/group_launches.csv \
        python profiles/group_probe.py 2 4000000 3
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))
import sphb200  # noqa: E402
from sphb200 import group as sg, ic  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
impl = sphb200.GRAVITY_TREE if (len(sys.argv) <= 4 or sys.argv[4] == "tree") else sphb200.GRAVITY_PARTICLE
c = ic.make_config("c4", particles=n)
g = sg.Group.single_process(n, [0] * world)
g.upload_global(c["pos"], c["vel"], c["mass"], c["h"])
g.enable_timing(True)
for k in range(steps):
    g.step(1 / 60, impl)
    print("step %d:" % k, {nm: round(ms, 3) for nm, ms in g.timings()}, flush=True)
g.sync()
print(g.info())
