"""Morton-range domain decomposition (sphb200_group_*) against the single-GPU handle.

The decomposition must not change a bit: with tree gravity every downloaded field of a group of any size equals the
single handle's (which tests/test_gpu_parity.py holds to the CPU oracle); with all-pairs gravity the source-split partial
sums depend on the number of targets per rank, so values agree to 1e-6 and everything else exactly.

On a single-GPU box the group's ranks share device 0 and exchange by in-process peer copies (same kernels, same
collective pattern, different transport); with >= 2 GPUs visible the same cases also run over NCCL (one process driving
all devices, and one process per GPU under torchrun).
Reference anchor: the decomposition replaces the reference's job-thread parallelism (UP/Collision/World/Broadphase.cs:163);
what it must reproduce is the single-handle step, i.e. SURVEY.md section 3.1.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FIELDS = ("pos", "vel", "h", "n_own", "rho", "P", "gradP", "grav", "count", "num_particles", "num_approx", "mass")


def _single(c, steps, impl, **params):
    import sphb200
    n = len(c["h"])
    sim = sphb200.Simulation(n, **params)
    sim.upload(c["pos"], c["vel"], c["mass"], c["h"])
    for _ in range(steps):
        sim.step(DT, impl)
    sim.sync()
    out = sim.download_all()
    diag = sim.diagnostics()
    sim.close()
    return out, diag


def _group(c, steps, impl, devices, **params):
    from sphb200 import group as sg
    n = len(c["h"])
    g = sg.Group.single_process(n, devices, **params)
    g.upload_global(c["pos"], c["vel"], c["mass"], c["h"])
    for _ in range(steps):
        g.step(DT, impl)
    g.sync()
    out = g.download_all()
    diag = g.diagnostics()
    info = g.info()
    g.close()
    return out, diag, info


def _compare(got, want, exact, tol=1e-6):
    bad = []
    for k in FIELDS:
        a, b = got[k], want[k]
        if exact or a.dtype.kind == "i" or k in ("pos", "h", "mass"):
            if not np.array_equal(a, b):
                bad.append("%s: %d of %d differ" % (k, int(np.sum(np.any(np.atleast_2d(a.T != b.T), axis=0))), len(a)))
        else:
            scale = np.abs(b).max() + 1e-30
            err = np.abs(a.astype(np.float64) - b).max() / scale
            if not err <= tol:
                bad.append("%s: rel err %.3g" % (k, err))
    assert not bad, "; ".join(bad)


def _sphere(n, seed=5, vel=0.3):
    from sphb200 import ic
    c = ic.make_sphere(n, radius=ic.scaled_radius(n), total_mass=100.0 * n / 3000, seed=seed)
    c["vel"] = np.random.default_rng(2).normal(0, vel, c["pos"].shape).astype(np.float32)
    return c


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_group_tree_gravity_bit_identical_to_single_handle(world):
    import sphb200
    c = _sphere(60013)                       # not divisible by any world size
    want, wd = _single(c, 4, sphb200.GRAVITY_TREE)
    got, gd, info = _group(c, 4, sphb200.GRAVITY_TREE, [0] * world)
    assert info["world"] == world and info["transport"] == ("none" if world == 1 else "local")
    assert sum(info["n_own"]) == len(c["h"])
    _compare(got, want, exact=True)
    assert gd["max_neighbors"] == wd["max_neighbors"] and abs(gd["mean_neighbors"] - wd["mean_neighbors"]) < 1e-9
    for k in ("mass", "e_kin", "e_pot", "e_int"):
        assert abs(gd[k] - wd[k]) <= 1e-9 * abs(wd[k]) + 1e-12, k


def test_group_locally_essential_tree_is_a_superset_of_what_the_walk_reads(monkeypatch):
    """Tree gravity of a group ships only the walk nodes another rank can reach (k_let_*): with every other node record
    poisoned (NaN moments, child -1) the step is still bit-identical to the single handle, far fewer records travel than the
    all-gather moved, and the all-gather fallback (SPHB200_GROUP_NO_LET) gives the same bits."""
    import sphb200
    c = _sphere(120011, seed=11)
    want, _ = _single(c, 3, sphb200.GRAVITY_TREE)
    monkeypatch.setenv("SPHB200_GROUP_LET_POISON", "1")
    got, _, info = _group(c, 3, sphb200.GRAVITY_TREE, [0, 0, 0, 0])
    _compare(got, want, exact=True)
    n = len(c["h"])
    # the all-gather moved (world-1)/world of 2n-1 records to every rank
    assert 0 < info["tree_nodes_last_step"] < 0.6 * 4 * (0.75 * 2 * n), info["tree_nodes_last_step"]
    monkeypatch.delenv("SPHB200_GROUP_LET_POISON")
    monkeypatch.setenv("SPHB200_GROUP_NO_LET", "1")
    got2, _, info2 = _group(c, 3, sphb200.GRAVITY_TREE, [0, 0, 0, 0])
    assert info2["tree_nodes_last_step"] == -1
    _compare(got2, want, exact=True)
    # collision geometry (two separated bodies, S > 1 stencil, unequal masses), poisoned
    monkeypatch.delenv("SPHB200_GROUP_NO_LET")
    monkeypatch.setenv("SPHB200_GROUP_LET_POISON", "1")
    from sphb200 import ic
    c5 = ic.make_collision(30000, seed=4)
    want5, _ = _single(c5, 2, sphb200.GRAVITY_TREE)
    got5, _, _ = _group(c5, 2, sphb200.GRAVITY_TREE, [0, 0, 0])
    _compare(got5, want5, exact=True)


@pytest.mark.parametrize("world", [2, 4])
def test_group_allpairs_gravity_matches_single_handle(world):
    import sphb200
    c = _sphere(30011)
    # one step: everything but the gravity sum (and the velocity it kicks) is bit-identical
    want, _ = _single(c, 1, sphb200.GRAVITY_PARTICLE)
    got, _, _ = _group(c, 1, sphb200.GRAVITY_PARTICLE, [0] * world)
    for k in ("pos", "h", "n_own", "count", "rho", "P", "gradP", "mass"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("grav", "vel"):
        assert np.abs(got[k].astype(np.float64) - want[k]).max() <= 1e-6 * np.abs(want[k]).max(), k
    # three steps: the 1e-6 differences of the kicks move positions by ulps, which may flip a neighbor at a support edge
    want, _ = _single(c, 3, sphb200.GRAVITY_PARTICLE)
    got, _, _ = _group(c, 3, sphb200.GRAVITY_PARTICLE, [0] * world)
    for k in ("pos", "vel", "h", "rho", "P", "gradP", "grav"):
        assert np.abs(got[k].astype(np.float64) - want[k]).max() <= 2e-5 * np.abs(want[k]).max(), k
    assert np.mean(got["count"] != want["count"]) < 1e-3


def test_group_collision_high_h_contrast_stencil_and_unequal_masses():
    """C5 geometry: 64x density contrast => stencil S > 1, unequal masses, two separated bodies (empty ranks' worth of
    cells between them), group velocity: halo selection and work-weighted splitters under stress."""
    import sphb200
    from sphb200 import ic
    c = ic.make_collision(20000, seed=3)
    want, _ = _single(c, 3, sphb200.GRAVITY_TREE)
    got, _, info = _group(c, 3, sphb200.GRAVITY_TREE, [0, 0, 0])
    _compare(got, want, exact=True)
    assert min(info["n_own"]) > 0


@pytest.mark.parametrize("leaf_max,aabb_mode", [(1, 0), (8, 1), (16, 0)])
def test_group_tree_variants(leaf_max, aabb_mode):
    import sphb200
    c = _sphere(20011, seed=9)
    want, _ = _single(c, 2, sphb200.GRAVITY_TREE, leaf_max=leaf_max, aabb_mode=aabb_mode)
    got, _, _ = _group(c, 2, sphb200.GRAVITY_TREE, [0, 0, 0], leaf_max=leaf_max, aabb_mode=aabb_mode)
    _compare(got, want, exact=True)


def test_group_tiny_and_empty_ranks():
    """Fewer particles than leaf buckets per rank: buckets straddle several rank boundaries, some ranks own nothing."""
    import sphb200
    for n in (1, 2, 7, 40, 333):
        c = _sphere(n, seed=n)
        want, _ = _single(c, 2, sphb200.GRAVITY_TREE)
        got, _, _ = _group(c, 2, sphb200.GRAVITY_TREE, [0, 0, 0, 0])
        _compare(got, want, exact=True)


def test_group_many_steps_migration_and_rebalancing():
    """20 steps with fast random motion: particles cross rank boundaries every step; the trajectory stays bit-identical."""
    import sphb200
    from sphb200 import group as sg
    c = _sphere(20000, seed=11, vel=3.0)
    n = len(c["h"])
    sim = sphb200.Simulation(n)
    sim.upload(c["pos"], c["vel"], c["mass"], c["h"])
    g = sg.Group.single_process(n, [0, 0, 0])
    g.upload_global(c["pos"], c["vel"], c["mass"], c["h"])
    migrated = 0
    for _ in range(20):
        sim.step(DT, sphb200.GRAVITY_TREE)
        g.step(DT, sphb200.GRAVITY_TREE)
        migrated += g.info()["migrated_last_step"]
    assert migrated > n            # the first step shuffles everything; later steps keep migrating
    _compare(g.download_all(), sim.download_all(), exact=True)
    sim.close(); g.close()


def test_group_component_struct_upload_download_and_strides():
    """The group takes the same component arrays / strides as the single handle (reference layouts)."""
    import sphb200
    from sphb200 import group as sg
    c = _sphere(5000, seed=4)
    n = len(c["h"])
    sm = np.zeros(n, sphb200.ParticleSmoothing); sm["influenceArea"] = c["h"]; sm["neighbors"] = 37
    pv = np.zeros(n, sphb200.PhysicsVelocity); pv["linear"] = c["vel"]
    sim = sphb200.Simulation(n); sim.upload(c["pos"], pv, c["mass"], sm); sim.step(DT, sphb200.GRAVITY_TREE)
    g = sg.Group.single_process(n, [0, 0]); g.upload(n, c["pos"], pv, c["mass"], sm); g.step(DT, sphb200.GRAVITY_TREE)
    _compare(g.download_all(), sim.download_all(), exact=True)
    wide = np.zeros((n, 5), np.float32)                      # a padded stride for a 3-float field
    g.download(sphb200.FIELD_PRESSURE_GRAD, wide)
    assert np.array_equal(wide[:, :3], sim.download(sphb200.FIELD_PRESSURE_GRAD)) and not wide[:, 3:].any()
    sim.close(); g.close()


def test_group_errors():
    import sphb200
    from sphb200 import group as sg
    g = sg.Group.single_process(1000, [0, 0])
    with pytest.raises(sphb200.SphError) as e:
        g.step(DT, sphb200.GRAVITY_TREE)
    assert e.value.code == sphb200.SPH_ERR_STATE
    c = _sphere(2000)
    with pytest.raises(sphb200.SphError) as e:
        g.upload_global(c["pos"], c["vel"], c["mass"], c["h"])
    assert e.value.code == sphb200.SPH_ERR_CAPACITY
    g.close()
    with pytest.raises(sphb200.SphError):
        sg.Group.single_process(1000, [0] * 40)


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_group_nccl_single_process_distinct_devices():
    """>= 2 GPUs: one process drives them over NCCL (ncclCommInitAll) -- the C# host's configuration."""
    import sphb200
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    c = _sphere(200003)
    want, _ = _single(c, 3, sphb200.GRAVITY_TREE)
    got, _, info = _group(c, 3, sphb200.GRAVITY_TREE, list(range(min(_ngpu(), 8))))
    assert info["transport"] == "nccl"
    _compare(got, want, exact=True)


def test_group_nccl_one_process_per_gpu_torchrun():
    """>= 2 GPUs: tests/mgpu_check.py under torchrun (sphb200_group_create_rank), every rank against a single handle."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
