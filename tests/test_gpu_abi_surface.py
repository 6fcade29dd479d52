"""GPU tests of the entry points of include/sphb200.h that no other test reaches: the gravity hint
(sphb200_prepare_gravity), running on a caller's stream (sphb200_set_stream / _get_stream), the zero-copy array getter
(sphb200_device_ptr) and the per-rank inspection handle of a group (sphb200_group_rank_handle).  Everything goes through
the C ABI (ctypes); results are compared bit for bit with the plain fused step, whose parity with the oracle is
tests/test_gpu_parity.py's subject."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DT = 0.02
FIELDS = ("pos", "vel", "h", "rho", "P", "gradP", "grav", "n_own", "num_particles", "num_approx")


def fused(c, impl, n=None):
    import sphb200
    sim = sphb200.Simulation(n or len(c["h"]))
    sim.upload(c["pos"], c["vel"], c["mass"], c["h"])
    sim.step(DT, impl)
    return sim


def assert_same(a, b):
    for k in FIELDS:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_prepare_gravity_hint_changes_the_schedule_not_the_bits():
    """The hint moves the LBVH build onto the auxiliary stream beside the neighbor pass (the reference builds its tree in
    BuildPhysicsWorld, before KernelSystem: BuildPhysicsWorld.cs:286-289).  Staged calls in the reference's system order, with
    and without the hint, with a hint for another dt (the tree is rebuilt) and with a hint that is then not used."""
    import sphb200
    from sphb200 import ic
    c = ic.make_sphere(6000, seed=21)
    want = fused(c, sphb200.GRAVITY_TREE).download_all()

    def staged(hint_impl, hint_dt, gravity_impl=sphb200.GRAVITY_TREE):
        sim = sphb200.Simulation(6000)
        sim.upload(c["pos"], c["vel"], c["mass"], c["h"])
        sim.smoothing_update()
        if hint_impl is not None:
            sim.prepare_gravity(hint_impl, hint_dt)
        sim.build_neighbors()
        sim.gravity(gravity_impl, DT)
        sim.density()
        sim.pressure()
        sim.integrate(DT)
        out = sim.download_all()
        sim.close()
        return out

    assert_same(staged(None, 0.0), want)
    assert_same(staged(sphb200.GRAVITY_TREE, DT), want)
    assert_same(staged(sphb200.GRAVITY_TREE, 0.5 * DT), want)          # MAC boxes are swept by v dt (quirk Q2): rebuilt for the dt used
    assert_same(staged(sphb200.GRAVITY_PARTICLE, DT), want)            # no tree announced: built inside gravity()
    direct = fused(c, sphb200.GRAVITY_PARTICLE).download_all()
    assert_same(staged(sphb200.GRAVITY_TREE, DT, sphb200.GRAVITY_PARTICLE), direct)   # announced but not used


def test_step_on_a_callers_stream_gives_the_same_bits():
    import torch
    import sphb200
    from sphb200 import ic
    c = ic.make_sphere(5000, seed=22)
    want = fused(c, sphb200.GRAVITY_TREE).download_all()
    sim = sphb200.Simulation(5000)
    own = sim.stream_ptr()
    assert own != 0
    s = torch.cuda.Stream()
    sim.set_stream(s.cuda_stream)
    assert sim.stream_ptr() == s.cuda_stream
    sim.upload(c["pos"], c["vel"], c["mass"], c["h"])
    sim.step(DT, sphb200.GRAVITY_TREE)
    # stream-ordered with the caller's own work: an event of that stream covers the step
    ev = torch.cuda.Event()
    ev.record(s)
    ev.synchronize()
    assert_same(sim.download_all(), want)
    sim.set_stream(0)                                                   # NULL = back to the handle's own stream
    assert sim.stream_ptr() == own
    sim.step(DT, sphb200.GRAVITY_TREE)
    two = fused(c, sphb200.GRAVITY_TREE)
    two.step(DT, sphb200.GRAVITY_TREE)
    assert_same(sim.download_all(), two.download_all())
    sim.close()


class _DeviceArray:
    """numba-style view of library-owned device memory for torch.as_tensor (zero copy)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def test_device_ptr_views_the_resident_sorted_arrays():
    import torch
    import sphb200
    from sphb200 import ic
    n, cap = 4000, 4500
    c = ic.make_sphere(n, seed=23)
    sim = fused(c, sphb200.GRAVITY_TREE, n=cap)
    got = sim.download_all()
    sim.sync()

    def view(name, typestr, width):
        ptr, nbytes = sim.device_ptr(name)
        assert ptr and nbytes == 4 * width * cap                       # sized by the capacity
        t = torch.as_tensor(_DeviceArray(ptr, n * width, typestr), device="cuda")
        return t.cpu().numpy().reshape(n, width) if width > 1 else t.cpu().numpy()

    orig = view("orig", "<i4", 1)                                       # sorted slot -> body index (uint32 in the library; < 2^31 here)
    assert np.array_equal(np.sort(orig), np.arange(n))
    np.testing.assert_array_equal(view("rho", "<f4", 1), got["rho"][orig])
    np.testing.assert_array_equal(view("press", "<f4", 1), got["P"][orig])
    np.testing.assert_array_equal(view("gradp", "<f4", 4)[:, :3], got["gradP"][orig])
    np.testing.assert_array_equal(view("grav", "<f4", 4), got["grav"][orig])
    posh = view("posh", "<f4", 4)
    np.testing.assert_array_equal(posh[:, :3], got["pos"][orig])
    np.testing.assert_array_equal(posh[:, 3], got["h"][orig])
    velm = view("velm", "<f4", 4)
    np.testing.assert_array_equal(velm[:, :3], got["vel"][orig])
    np.testing.assert_array_equal(velm[:, 3], c["mass"][orig])
    np.testing.assert_array_equal(view("nown", "<i4", 1), got["n_own"][orig])
    with pytest.raises(sphb200.SphError) as e:
        sim.device_ptr("no_such_array")
    assert e.value.code == sphb200.SPH_ERR_INVALID_ARG
    sim.close()


def test_group_rank_handle_exposes_each_ranks_context():
    import sphb200
    from sphb200 import ic
    from sphb200.group import Group
    n = 9000
    c = ic.make_sphere(n, seed=24)
    g = Group.single_process(n, [0, 0, 0])                              # three ranks on one device: in-process transport
    g.upload_global(c["pos"], c["vel"], c["mass"], c["h"])
    g.step(DT, sphb200.GRAVITY_TREE)
    g.sync()
    info = g.info()
    assert info["world"] == 3 and info["nlocal"] == 3 and sum(info["n_own"]) == n
    L = g.L
    for k in range(3):
        h = C.c_void_p()
        assert L.sphb200_group_rank_handle(g.h, k, C.byref(h)) == sphb200.SPH_OK and h.value
        cnt, cap = C.c_int64(), C.c_int64()
        assert L.sphb200_count(h, C.byref(cnt), C.byref(cap)) == sphb200.SPH_OK
        assert cnt.value == info["n_own"][k] + info["n_halo"][k] <= cap.value     # resident = [low halo | own | high halo]
        p = sphb200.Params()
        assert L.sphb200_get_params(h, C.byref(p)) == sphb200.SPH_OK
        assert p.leaf_max == g.params.leaf_max and p.max_neighbors == g.params.max_neighbors and p.theta == g.params.theta
        s = C.c_void_p()
        assert L.sphb200_get_stream(h, C.byref(s)) == sphb200.SPH_OK and (s.value or 0) == g.stream_ptr(k)
    bad = C.c_void_p()
    assert L.sphb200_group_rank_handle(g.h, 3, C.byref(bad)) == sphb200.SPH_ERR_INVALID_ARG
    assert L.sphb200_group_rank_handle(g.h, -1, C.byref(bad)) == sphb200.SPH_ERR_INVALID_ARG
    # the group's step equals the single handle's, bit for bit (tree gravity)
    want = fused(c, sphb200.GRAVITY_TREE).download_all()
    got = g.download_all()
    for k in ("pos", "vel", "rho", "P", "gradP", "grav"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    g.close()
