"""Host-side statement of the Morton-range decomposition rules that csrc/kernels_group.cu implements on the device
(numpy, no GPU): body slices, key-prefix bins, work-balanced splitters, owner lookup and halo destination masks.

Used by the CPU tests (tests/test_host_sharding.py runs a two-rank decomposition over gloo with these rules and checks the
two claims the GPU path relies on: rank-ordered stable local sorts reproduce the global stable sort, and the halo is a
superset of every neighbor) and by host code that wants to know which rank owns what.  Nothing here is on the hot path.
"""
import numpy as np

BIN_BITS = 18                      # SPH_BIN_BITS (csrc/group.cuh)


def body_slice(n, rank, world):
    """Bodies [b0, b1) that `rank` uploads / downloads (sphb200_group_body_range for a one-rank process)."""
    chunk = max((int(n) + world - 1) // world, 1)
    b0 = min(rank * chunk, int(n))
    return b0, min(b0 + chunk, int(n))


def bin_bits(grid_bits):
    return min(BIN_BITS, 3 * int(grid_bits))


def bins_of(keys, grid_bits):
    """Ownership bin of 30-bit Morton keys: a key prefix, never finer than a grid cell."""
    return np.asarray(keys, np.uint32) >> np.uint32(30 - bin_bits(grid_bits))


def work_of(ncount, npart, napprox):
    """Balance weight of a particle (k_bin_hist): 8 per neighbor (+1) and 1 per tree interaction of the last step."""
    return 8 * (np.maximum(ncount, 0).astype(np.int64) + 1) + np.maximum(npart, 0) + np.maximum(napprox, 0)


def splitters(count_hist, work_hist, world):
    """k_splitters: rank r owns bins [sbin[r], sbin[r+1]) and global sorted slots [g0[r], g0[r+1]); splitter r is the
    first bin whose exclusive work prefix reaches r/world of the total work."""
    count_hist = np.asarray(count_hist, np.int64); work_hist = np.asarray(work_hist, np.int64)
    nb = len(count_hist)
    pw = np.concatenate([[0], np.cumsum(work_hist)])         # pw[b] = work of bins < b
    pc = np.concatenate([[0], np.cumsum(count_hist)])
    tot = int(pw[-1])
    g0 = np.zeros(world + 1, np.int64); sbin = np.zeros(world + 1, np.int64)
    g0[world] = pc[-1]; sbin[world] = nb
    for r in range(1, world):
        target = tot // world * r + (tot % world) * r // world
        if target > 0:
            b = int(np.searchsorted(pw, target, side="left"))  # first b with pw[b] >= target
            sbin[r] = b; g0[r] = pc[b]
    return g0, sbin


def owner_of_bins(bins, sbin):
    """Rank that owns each bin (k_dest): number of splitters 1..world-1 that are <= bin."""
    world = len(sbin) - 1
    return np.searchsorted(np.asarray(sbin[1:world], np.int64), np.asarray(bins, np.int64), side="right")


def _expand10(v):
    v = v.astype(np.uint32) & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def _compact10(v):
    v = v.astype(np.uint32) & 0x09249249
    v = (v | (v >> 2)) & 0x030C30C3
    v = (v | (v >> 4)) & 0x0300F00F
    v = (v | (v >> 8)) & 0x030000FF
    v = (v | (v >> 16)) & 0x3FF
    return v


def halo_mask(keys, grid_bits, stencil, sbin, me):
    """k_halo_mask: bit q set <=> some cell within `stencil` cells of the particle's cell belongs to rank q != me."""
    keys = np.asarray(keys, np.uint32)
    world = len(sbin) - 1
    ck = keys >> np.uint32(3 * (10 - grid_bits))
    cx, cy, cz = _compact10(ck).astype(np.int64), _compact10(ck >> 1).astype(np.int64), _compact10(ck >> 2).astype(np.int64)
    dim = 1 << grid_bits
    bshift = 3 * grid_bits - bin_bits(grid_bits)
    mask = np.zeros(len(keys), np.uint32)
    S = int(stencil)
    for oz in range(-S, S + 1):
        for oy in range(-S, S + 1):
            for ox in range(-S, S + 1):
                nx, ny, nz = cx + ox, cy + oy, cz + oz
                ok = (nx >= 0) & (ny >= 0) & (nz >= 0) & (nx < dim) & (ny < dim) & (nz < dim)
                nk = _expand10(np.clip(nx, 0, dim - 1)) | (_expand10(np.clip(ny, 0, dim - 1)) << 1) | (_expand10(np.clip(nz, 0, dim - 1)) << 2)
                r = owner_of_bins(nk >> np.uint32(bshift), sbin)
                mask |= np.where(ok, np.uint32(1) << r.astype(np.uint32), np.uint32(0))
    return mask & ~np.uint32(1 << me) & np.uint32((1 << world) - 1)


def walk_box(pos_sorted, g0, rank):
    """Box of the positions rank `rank` walks (k_let_box): its targets [g0[r], g0[r+1]) plus the companions that complete its
    first and last 32-slot group.  (lo, hi) float32[3]; an empty rank gets lo = +inf."""
    n = len(pos_sorted)
    a, b = int(g0[rank]), int(g0[rank + 1])
    if b <= a:
        return np.full(3, np.inf, np.float32), np.full(3, -np.inf, np.float32)
    lo, hi = a & ~31, min(n, (b + 31) & ~31)
    p = np.asarray(pos_sorted, np.float32)[lo:hi]
    return p.min(axis=0), p.max(axis=0)


def let_mask(first, last, parent, cm, b_sq, theta, g0, me, boxes, leaf_max):
    """k_let_mask: for every LBVH node (2n-1; arrays of the oracle's tree) a bit mask of the ranks other than `me` that may read
    the node's walk record, for the nodes rank `me` finishes (particle range inside [g0[me], g0[me+1])).  A walk reads node k only
    after one of its targets has opened k's parent p, i.e. NOT accepted it: RN(b_sq[p] / r_sq) >= theta^2.  r_sq is at least the
    squared distance from cm[p] to the box of the walking rank's positions, so "dist^2(cm[p], box_r) * 0.99999 <= T[p]" with
    T = b_sq / theta^2 (the device uses the exact ulp-stepped threshold; 1e-5 of slack covers it) is a superset of "some target
    of rank r opens p".  Nodes strictly inside a bucket are unreachable; nodes whose parent straddles a rank boundary (or is the
    root of another rank's subtree) are needed by everybody."""
    first = np.asarray(first, np.int64); last = np.asarray(last, np.int64); parent = np.asarray(parent, np.int64)
    nn = len(first); world = len(g0) - 1
    a, b = int(g0[me]), int(g0[me + 1])
    everyone = ((1 << world) - 1) & ~(1 << me)
    mask = np.zeros(nn, np.int64)
    size = last - first + 1
    theta2 = np.float32(theta) * np.float32(theta)
    for k in range(nn):
        if first[k] < a or last[k] >= b:
            continue                                   # not finished by this rank (another rank's node, or a straddling top node)
        p = parent[k]
        if p < 0:
            mask[k] = everyone
            continue
        if size[p] <= leaf_max:
            continue                                   # strictly inside a bucket
        if first[p] < a or last[p] >= b:
            mask[k] = everyone                         # the parent is a top node
            continue
        T = np.float32(b_sq[p]) / theta2
        m = 0
        for r in range(world):
            if r == me:
                continue
            lo, hi = boxes[r]
            d = np.maximum(np.maximum(lo - cm[p], cm[p] - hi), 0).astype(np.float32)
            if float(np.dot(d, d)) * 0.99999 <= float(T):
                m |= 1 << r
        mask[k] = m
    return mask
