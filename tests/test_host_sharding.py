"""CPU tests of the multi-GPU host logic: Morton-range partition arithmetic and the in-place slice
all-gather, exercised with the gloo backend at world_size 2 -- with the CPU oracle standing in for the GPU kernels, the
sharded step must reproduce the single-process oracle step exactly."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_arithmetic():
    from sphb200 import dist
    for n in (1, 2, 7, 1000, 200003, 16_000_000):
        for world in (1, 2, 3, 4, 8):
            ranges = [dist.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))          # contiguous, disjoint cover
            assert all(0 <= t1 - t0 <= dist.chunk_size(n, world) for t0, t1 in ranges)
            assert dist.padded_capacity(n, world) >= n and dist.padded_capacity(n, world) % world == 0
            assert max(t1 - t0 for t0, t1 in ranges) - min(t1 - t0 for t0, t1 in ranges[:-1] or ranges) <= dist.chunk_size(n, world)


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))
    import torch
    import torch.distributed as td
    from sphb200 import dist, ic
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    c = ic.make_sphere(n, seed=3)
    cap = dist.padded_capacity(n, world)
    t0, t1 = dist.shard_range(n, rank, world)
    # "sorted order" of the oracle = key order; every rank derives it from identical inputs
    g = orc.grid_params(c["pos"], c["h"], 5)
    order = orc.sort_order(orc.morton_keys(c["pos"], g)).astype(np.int64)
    pos, h, m = c["pos"][order], c["h"][order], c["mass"][order]
    off, nbr = orc.neighbors(pos, h)
    # stage 1 (own targets only): density -> cvol slice, then the "halo" exchange = slice all-gather
    rho_all, _ = orc.density(pos, h, m, off, nbr)
    P_all = orc.eos(rho_all)
    cvol = torch.zeros(cap, 1)
    cvol[t0:t1, 0] = torch.from_numpy((m / rho_all * P_all)[t0:t1])
    dist.allgather_slices(cvol, rank, world)
    # stage 2 (own targets): pressure gradient from the gathered cvol, written into the rank's slice, gathered again
    gp_all = orc.pressure_grad(pos, h, m, rho_all, P_all, off, nbr)
    want_cvol = (m / rho_all * P_all).astype(np.float32)
    ok = np.array_equal(cvol[:n, 0].numpy(), want_cvol)
    gp = torch.zeros(cap, 3)
    gp[t0:t1] = torch.from_numpy(gp_all[t0:t1])
    dist.allgather_slices(gp, rank, world)
    ok = ok and np.array_equal(gp[:n].numpy(), gp_all)
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([ok, t0, t1]))
    td.destroy_process_group()


def test_slice_allgather_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    n = 1501                                   # odd: the last rank owns one slot fewer, padding is exercised
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "ok0.npy"); r1 = np.load(tmp_path / "ok1.npy")
    assert r0[0] == 1 and r1[0] == 1
    assert (r0[1], r0[2], r1[1], r1[2]) == (0, 751, 751, 1501)
