"""Multi-GPU consistency check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank runs the Morton-range-sharded step (sphb200.dist.ShardedSimulation); rank 0 also runs the same steps on a
plain single-GPU Simulation and the two must agree: bit-for-bit with tree gravity (same kernels, same inputs, targets
merely partitioned), to 1e-6 with all-pairs gravity (the source-split partial sums depend on the target count).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))

import sphb200  # noqa: E402
from sphb200 import dist as sdist, ic  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(os.environ.get("MGPU_N", "200003"))          # deliberately not divisible by the world size
    c = ic.make_sphere(n, radius=ic.scaled_radius(n), total_mass=100.0 * n / 3000, seed=5)
    c["vel"] = np.random.default_rng(2).normal(0, 0.3, c["pos"].shape).astype(np.float32)
    ok = True
    for impl, name, tol in ((sphb200.GRAVITY_TREE, "tree", 0.0), (sphb200.GRAVITY_PARTICLE, "direct", 1e-6)):
        eng = sdist.ShardedSimulation(n, device=local, rank=rank, world=world)
        eng.upload(c["pos"], c["vel"], c["mass"], c["h"])
        for _ in range(4):
            eng.step(1 / 60, impl)
        eng.gather_results()
        eng.sim.sync()
        got = eng.sim.download_all()
        if rank == 0:
            ref = sphb200.Simulation(n, device=local)
            ref.upload(c["pos"], c["vel"], c["mass"], c["h"])
            for _ in range(4):
                ref.step(1 / 60, impl)
            want = ref.download_all()
            for k in ("pos", "vel", "h", "n_own", "rho", "P", "gradP", "grav", "count"):
                a, b = got[k], want[k]
                if tol == 0.0 or a.dtype.kind == "i":
                    same = np.array_equal(a, b)
                else:
                    scale = np.abs(b).max() + 1e-30
                    same = np.abs(a.astype(np.float64) - b).max() <= tol * scale
                print("[mgpu] %s %-6s %s" % (name, k, "ok" if same else "MISMATCH"))
                ok = ok and same
        td.barrier()
    if rank == 0:
        print("[mgpu] world=%d n=%d %s" % (world, n, "ALL OK" if ok else "FAILED"))
    td.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
