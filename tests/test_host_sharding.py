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


def test_locally_essential_tree_rule_covers_every_node_the_oracle_walk_reads():
    """The sender-side box test of the group's tree exchange (decomp.let_mask = k_let_mask) is a superset of what the walks of
    the other ranks read: every node the oracle's per-particle walk (orc_tree_walk semantics, AcceptApproximation with the
    reference's fp32 arithmetic) pops for a target of rank r, and that another rank finished, carries bit r in that rank's mask --
    for a uniform sphere and for the two-body collision geometry (64x density contrast), 3 and 4 ranks, unequal slot ranges."""
    from oracle import oracle as orc
    from sphb200 import decomp, ic
    f32 = np.float32
    for c, world, cuts in ((ic.make_sphere(1500, seed=4), 3, (0.0, 0.22, 0.71, 1.0)),
                           (ic.make_collision(700, seed=5), 4, (0.0, 0.3, 0.5, 0.83, 1.0))):
        pos, vel, h, m = (np.asarray(c[k], np.float32) for k in ("pos", "vel", "h", "mass"))
        n = len(h)
        g = orc.grid_params(pos, h, 5)
        keys = orc.morton_keys(pos, g)
        order = orc.sort_order(keys).astype(np.int64)
        ps, vs, hs, ms = pos[order], vel[order], h[order], m[order]
        leaf_max, theta, dt = 4, 0.7, 1 / 60
        t = orc.lbvh_build(keys[order], ps, vs, hs, ms, leaf_max, 0, dt)
        cm = t.mom[:, :3]
        bx = np.maximum(t.hi - cm, cm - t.lo).astype(f32)
        b_sq = ((bx[:, 0] * bx[:, 0] + bx[:, 1] * bx[:, 1]).astype(f32) + bx[:, 2] * bx[:, 2]).astype(f32)    # AcceptApproximation order
        theta2 = f32(theta) * f32(theta)
        g0 = np.array([int(round(x * n)) for x in cuts], np.int64)
        boxes = [decomp.walk_box(ps, g0, r) for r in range(world)]
        masks = [decomp.let_mask(t.first, t.last, t.parent, cm, b_sq, theta, g0, r, boxes, leaf_max) for r in range(world)]
        finished_by = np.full(2 * n - 1, -1)
        for r in range(world):
            inside = (t.first >= g0[r]) & (t.last < g0[r + 1])
            finished_by[inside] = r
        size = t.last - t.first + 1
        read = 0
        for r in range(world):
            # the walking slots of rank r: its targets and the companions of its first / last 32-slot group
            lo, hi = int(g0[r]) & ~31, min(n, (int(g0[r + 1]) + 31) & ~31)
            for s in range(lo, hi):
                stack = [0]
                while stack:
                    k = stack.pop()
                    owner = finished_by[k]
                    if owner >= 0 and owner != r:
                        read += 1
                        assert (masks[owner][k] >> r) & 1, "rank %d reads node %d of rank %d outside its essential tree" % (r, k, owner)
                    d = (ps[s] - cm[k]).astype(f32)
                    r_sq = ((d[0] * d[0] + d[1] * d[1]).astype(f32) + d[2] * d[2]).astype(f32)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        accept = (b_sq[k] / r_sq) < theta2
                    if accept or size[k] <= leaf_max:
                        continue
                    stack.append(int(t.left[k])); stack.append(int(t.right[k]))
        assert read > 0
        # and it is selective: far fewer records than the whole remote tree
        sent = sum(int(np.count_nonzero(masks[o] & (1 << r))) for o in range(world) for r in range(world) if r != o)
        assert sent < 0.9 * (world - 1) * (2 * n - 1)
