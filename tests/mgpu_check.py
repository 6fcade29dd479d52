"""Multi-GPU consistency check, one process per GPU (run under torchrun on a box with >= 2 GPUs; tests/test_group.py launches
it as a gpu-marked pytest case when 2 or more GPUs are visible):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank joins a Morton-range-decomposed group (sphb200_group_create_rank: NCCL over NVLink), uploads its body slice,
steps, downloads its slice; it also runs the same steps on a plain single-GPU handle and the two must agree: bit for bit
with tree gravity, to 1e-6 with all-pairs gravity (the source-split partial sums depend on the targets per rank).
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
from sphb200 import group as sgroup, ic  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("gloo")                        # carries only the 128-byte NCCL id; the data path is the library's NCCL
    n = int(os.environ.get("MGPU_N", "200003"))          # deliberately not divisible by the world size
    steps = int(os.environ.get("MGPU_STEPS", "4"))
    c = ic.make_sphere(n, radius=ic.scaled_radius(n), total_mass=100.0 * n / 3000, seed=5)
    c["vel"] = np.random.default_rng(2).normal(0, 0.3, c["pos"].shape).astype(np.float32)
    ok = True
    for impl, name, tol in ((sphb200.GRAVITY_TREE, "tree", 0.0), (sphb200.GRAVITY_PARTICLE, "direct", 1e-6)):
        g = sgroup.Group.from_env(n, device=local)
        g.upload_global(c["pos"], c["vel"], c["mass"], c["h"])
        for _ in range(steps):
            g.step(1 / 60, impl)
        g.sync()
        got = g.download_all()
        info = g.info()
        b0, cnt = g.body0, g.count
        ref = sphb200.Simulation(n, device=local)
        ref.upload(c["pos"], c["vel"], c["mass"], c["h"])
        for _ in range(steps):
            ref.step(1 / 60, impl)
        want = ref.download_all()
        ref.close()
        for k in ("pos", "vel", "h", "n_own", "rho", "P", "gradP", "grav", "count", "num_particles", "num_approx"):
            a, b = got[k], want[k][b0:b0 + cnt]
            if tol == 0.0 or a.dtype.kind == "i":
                same = np.array_equal(a, b)
            else:
                scale = np.abs(want[k]).max() + 1e-30
                same = np.abs(a.astype(np.float64) - b).max() <= tol * scale
            if not same:
                print("[mgpu] rank %d %s %-6s MISMATCH" % (rank, name, k))
            ok = ok and same
        print("[mgpu] rank %d %s: own %s halo %s transport %s %s" % (rank, name, info["n_own"], info["n_halo"], info["transport"],
                                                                      "ok" if ok else "FAILED"))
        g.close()
        td.barrier()
    flag = torch.tensor([0 if ok else 1])
    td.all_reduce(flag)
    if rank == 0:
        print("[mgpu] world=%d n=%d %s" % (world, n, "ALL OK" if flag.item() == 0 else "FAILED"))
    td.destroy_process_group()
    if flag.item() != 0:
        sys.exit(1)


if __name__ == "__main__":
    main()
