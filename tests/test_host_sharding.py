"""CPU tests of the multi-GPU host logic: a two-rank Morton-range decomposition run over gloo with the rules
of sphb200/decomp.py (the numpy statement of csrc/kernels_group.cu), the CPU oracle standing in for the GPU kernels.

Checked: (1) the body-slice arithmetic; (2) the ordering claim -- every rank stably sorts what it received in source-rank
order, and the concatenation over ranks IS the global stable sort (oracle: orc_sort_order), including duplicate keys;
(3) the halo claim -- own + halo particles contain every oracle neighbor of every own particle, also with a stencil S > 1;
(4) the 128-byte id broadcast that Group.from_env uses."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_body_slices_cover_disjointly():
    from sphb200 import decomp
    for n in (0, 1, 2, 7, 1000, 200003, 16_000_000):
        for world in (1, 2, 3, 4, 8, 32):
            s = [decomp.body_slice(n, r, world) for r in range(world)]
            assert s[0][0] == 0 and s[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(s, s[1:]))
            assert max(b - a for a, b in s) <= (n + world - 1) // world or n == 0


def test_splitters_balance_work_and_respect_bins():
    from sphb200 import decomp
    rng = np.random.default_rng(0)
    cnt = rng.integers(0, 50, 4096); work = cnt * rng.integers(8, 3000, 4096)
    for world in (1, 2, 3, 8, 32):
        g0, sbin = decomp.splitters(cnt, work, world)
        assert g0[0] == 0 and g0[-1] == cnt.sum() and sbin[0] == 0 and sbin[-1] == 4096
        assert np.all(np.diff(g0) >= 0) and np.all(np.diff(sbin) >= 0)
        per = [work[sbin[r]:sbin[r + 1]].sum() for r in range(world)]
        assert max(per) <= work.sum() / world + work.max()          # never more than one bin above the even share
        owner = decomp.owner_of_bins(np.arange(4096), sbin)
        assert all(np.all(owner[sbin[r]:sbin[r + 1]] == r) for r in range(world))
    g0, sbin = decomp.splitters(np.zeros(64, int), np.zeros(64, int), 4)   # no particles at all
    assert not g0.any() and list(sbin) == [0, 0, 0, 0, 64]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))
    import torch
    import torch.distributed as td
    from sphb200 import decomp, ic
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for case, bits in (("sphere", 4), ("collision", 5)):
        if case == "sphere":
            c = ic.make_sphere(3001, seed=3)
            c["pos"][10:40] = c["pos"][10]                        # coincident particles: duplicate keys across the body slices
        else:
            c = ic.make_collision(1500, seed=2)                   # 64x density contrast: stencil S > 1
        n = len(c["h"])
        g = orc.grid_params(c["pos"], c["h"], bits)               # identical on every rank (device: all-reduced bounds)
        keys = orc.morton_keys(c["pos"], g)
        b0, b1 = decomp.body_slice(n, rank, world)
        mine = np.arange(b0, b1)                                  # previous order = body order, split in body slices
        # ownership: all-reduced histograms -> splitters -> destination of every own particle
        nb = 1 << decomp.bin_bits(g.bits)
        bins = decomp.bins_of(keys[mine], g.bits)
        cnt = torch.from_numpy(np.bincount(bins, minlength=nb).astype(np.int64))
        work = torch.from_numpy(np.bincount(bins, weights=np.full(len(mine), 8.0), minlength=nb).astype(np.int64))
        td.all_reduce(cnt); td.all_reduce(work)
        g0, sbin = decomp.splitters(cnt.numpy(), work.numpy(), world)
        dest = decomp.owner_of_bins(bins, sbin)
        # migration: stable per-destination order, received in source-rank order
        send = [mine[dest == q] for q in range(world)]
        recv = [None] * world
        td.all_gather_object(recv, send)
        got = np.concatenate([recv[p][rank] for p in range(world)])
        local = got[np.argsort(keys[got], kind="stable")]         # the rank's stable radix sort
        allr = [None] * world
        td.all_gather_object(allr, local)
        order = orc.sort_order(keys).astype(np.int64)             # the single-GPU order
        ok = ok and np.array_equal(np.concatenate(allr), order)
        ok = ok and len(local) == g0[rank + 1] - g0[rank]
        # halo: destinations of my sorted particles; what I receive must cover every neighbor of my particles
        mask = decomp.halo_mask(keys[local], g.bits, g.stencil, sbin, rank)
        hsend = [local[(mask >> q) & 1 == 1] for q in range(world)]
        hrecv = [None] * world
        td.all_gather_object(hrecv, hsend)
        have = np.zeros(n, bool)
        have[local] = True
        for p in range(world):
            have[hrecv[p][rank]] = True
        off, nbr = orc.neighbors(c["pos"], c["h"])
        for i in local:
            ok = ok and bool(np.all(have[nbr[off[i]:off[i + 1]]]))
        ok = ok and (case != "collision" or g.stencil > 1)
    # the id broadcast of Group.from_env (128 bytes from rank 0)
    t = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    td.broadcast(t, 0)
    ok = ok and t.tolist() == list(range(128))
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([ok]))
    td.destroy_process_group()


def test_two_rank_decomposition_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "ok0.npy")[0] and np.load(tmp_path / "ok1.npy")[0]
