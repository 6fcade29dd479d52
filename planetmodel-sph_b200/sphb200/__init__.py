"""sphb200 -- Python host mirror of the PlanetModel-SPH per-timestep systems over the libsphb200 C ABI.

The product is the CUDA library (``libsphb200.so``, built in-tree from ``../csrc`` for sm_100a); this package is the
thin host side a test driver or benchmark uses in place of Unity's ECS world.  There is NO CPU fallback: importing
works without a GPU (so that ABI/export checks can run), but creating a :class:`Simulation` fails loudly when the
library or a B200 is missing.  Nothing here imports ``oracle/``.

Reference interface mirrored: the six SystemBase classes of Assets/Scripts/Systems and their
update order inside FixedStepSimulationSystemGroup -- see :mod:`sphb200.systems`.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsphb200.so")

SPH_OK = 0
SPH_ERR_INVALID_ARG = -1
SPH_ERR_CAPACITY = -2
SPH_ERR_NEIGHBOR_OVERFLOW = -3
SPH_ERR_CUDA = -4
SPH_ERR_STATE = -5
SPH_ERR_TREE_STACK = -6
SPH_ERR_NCCL = -7

GRAVITY_TREE, GRAVITY_PARTICLE, GRAVITY_NONE = 0, 1, 2
FLAG_FIX_KERNEL_DERIV_SIGN = 1
FLAG_KICK_DRIFT = 2
FLAG_PM07_SOFTENING = 4     # Price & Monaghan 2007 spline-softened, h-symmetric gravity (reference roadmap README.md:75-77)

(FIELD_TRANSLATION, FIELD_VELOCITY, FIELD_MASS, FIELD_SMOOTHING, FIELD_DENSITY, FIELD_PRESSURE, FIELD_PRESSURE_GRAD,
 FIELD_GRAVITY, FIELD_NEIGHBOR_COUNT) = range(9)


class Params(C.Structure):
    _fields_ = [("K", C.c_float), ("G", C.c_float), ("theta", C.c_float), ("target_neighbors", C.c_float),
                ("max_neighbors", C.c_int32), ("leaf_max", C.c_int32), ("aabb_mode", C.c_int32),
                ("max_grid_bits", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32 * 3)]


class GridParams(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("cell", C.c_float), ("fine_scale", C.c_float), ("bits", C.c_int32),
                ("hmax", C.c_float), ("ext", C.c_float), ("href", C.c_float), ("stencil", C.c_int32)]


# byte-exact mirrors of the reference components (include/sphb200.h)
Translation = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4")])
PhysicsVelocity = np.dtype([("linear", "f4", 3), ("angular", "f4", 3)])
ParticleSmoothing = np.dtype([("influenceArea", "f4"), ("supportDomain", "f4"), ("sphereColliderPosRadius", "f4", 4),
                              ("neighbors", "i4")])
GravityField = np.dtype([("value", "f4", 4), ("numParticles", "i4"), ("numApprox", "i4")])
ParticleInteraction = np.dtype([("otherIndex", "i4"), ("otherVersion", "i4"), ("kernelThis", "f4", 4),
                                ("kernelSymmetric", "f4", 4)])
assert Translation.itemsize == 12 and PhysicsVelocity.itemsize == 24 and ParticleSmoothing.itemsize == 28
assert GravityField.itemsize == 24 and ParticleInteraction.itemsize == 40

EXPORTS = [
    "sphb200_default_params", "sphb200_create", "sphb200_destroy", "sphb200_last_error", "sphb200_set_stream",
    "sphb200_sync", "sphb200_upload", "sphb200_smoothing_update", "sphb200_build_neighbors", "sphb200_gravity",
    "sphb200_prepare_gravity", "sphb200_density", "sphb200_pressure", "sphb200_integrate", "sphb200_step", "sphb200_set_target_range",
    "sphb200_download", "sphb200_download_neighbors", "sphb200_download_interactions", "sphb200_download_sort",
    "sphb200_download_tree", "sphb200_diagnostics", "sphb200_count", "sphb200_get_params", "sphb200_device_ptr",
    "sphb200_launch_count", "sphb200_enable_timing", "sphb200_get_timings", "sphb200_fp32_peak", "sphb200_version",
    "sphb200_group_create", "sphb200_group_unique_id", "sphb200_group_create_rank", "sphb200_group_destroy",
    "sphb200_group_last_error", "sphb200_group_body_range", "sphb200_group_upload", "sphb200_group_step",
    "sphb200_group_download", "sphb200_group_sync", "sphb200_group_diagnostics", "sphb200_group_info",
    "sphb200_group_enable_timing", "sphb200_group_get_timings", "sphb200_group_rank_handle", "sphb200_group_stream",
    "sphb200_get_stream", "sphb200_field_stats", "sphb200_snapshot_save", "sphb200_snapshot_load",
    "sphb200_group_field_stats", "sphb200_group_snapshot_save", "sphb200_group_snapshot_load", "sphb200_debug_check_guards",
]


class GroupInfo(C.Structure):
    _fields_ = [("world", C.c_int32), ("nlocal", C.c_int32), ("rank0", C.c_int32), ("transport", C.c_int32),
                ("n_total", C.c_int64), ("steps", C.c_int64), ("migrated_last_step", C.c_int64), ("halo_last_step", C.c_int64),
                ("cap_own", C.c_int64), ("cap_halo", C.c_int64), ("launches", C.c_int64), ("tree_nodes_last_step", C.c_int64),
                ("n_own", C.c_int64 * 32), ("n_halo", C.c_int64 * 32)]


def debug_check_guards():
    """(bad guard bytes, live allocations) of the SPH_DEBUG_BOUNDS build of the library; allocations == -1: release build."""
    L = load_library()
    bad = C.c_int64(0); na = C.c_int64(0)
    rc = L.sphb200_debug_check_guards(C.byref(bad), C.byref(na))
    if rc:
        raise RuntimeError("sphb200_debug_check_guards failed: %d" % rc)
    return int(bad.value), int(na.value)


class SphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sphb200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library():
    """dlopen libsphb200.so (no GPU needed to load it). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # fresh checkout (the .so is git-ignored): build it in-tree when a CUDA toolchain is present
        import shutil
        import subprocess
        nvcc = shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)
        if nvcc:
            subprocess.call(["make", "-C", os.path.join(os.path.dirname(_HERE), "csrc"), "-j4"], stdout=subprocess.DEVNULL)
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsphb200.so is not built (run `python __graft_entry__.py` or `make -C planetmodel-sph_b200/csrc`); "
                          "there is no CPU fallback")
    L = C.CDLL(os.environ.get("SPHB200_LIB") or LIB_PATH)    # SPHB200_LIB: another build of the same library (A/B timing of kernel variants)
    H = C.c_void_p
    L.sphb200_default_params.argtypes = [C.POINTER(Params)]
    L.sphb200_create.argtypes = [C.POINTER(Params), C.c_int64, C.c_int, C.POINTER(H)]
    L.sphb200_destroy.argtypes = [H]
    L.sphb200_last_error.argtypes = [H]
    L.sphb200_last_error.restype = C.c_char_p
    L.sphb200_set_stream.argtypes = [H, C.c_void_p]
    L.sphb200_sync.argtypes = [H]
    L.sphb200_upload.argtypes = [H, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    for fn in (L.sphb200_smoothing_update, L.sphb200_build_neighbors, L.sphb200_density, L.sphb200_pressure):
        fn.argtypes = [H]
    L.sphb200_gravity.argtypes = [H, C.c_int, C.c_float]
    L.sphb200_prepare_gravity.argtypes = [H, C.c_int, C.c_float]
    L.sphb200_integrate.argtypes = [H, C.c_float]
    L.sphb200_step.argtypes = [H, C.c_float, C.c_int]
    L.sphb200_set_target_range.argtypes = [H, C.c_int64, C.c_int64]
    L.sphb200_download.argtypes = [H, C.c_int, C.c_void_p, C.c_int]
    L.sphb200_download_neighbors.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.sphb200_download_interactions.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphb200_download_sort.argtypes = [H, C.c_void_p, C.c_void_p, C.POINTER(GridParams)]
    L.sphb200_download_tree.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphb200_diagnostics.argtypes = [H, C.c_void_p]
    L.sphb200_count.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.sphb200_get_params.argtypes = [H, C.POINTER(Params)]
    L.sphb200_device_ptr.argtypes = [H, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.sphb200_launch_count.argtypes = [H, C.POINTER(C.c_int64)]
    L.sphb200_enable_timing.argtypes = [H, C.c_int]
    L.sphb200_get_timings.argtypes = [H, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.sphb200_fp32_peak.argtypes = [H, C.POINTER(C.c_double)]
    L.sphb200_version.restype = C.c_char_p
    L.sphb200_debug_check_guards.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.sphb200_group_create.argtypes = [C.POINTER(Params), C.c_int64, C.c_int, C.POINTER(C.c_int), C.POINTER(H)]
    L.sphb200_group_unique_id.argtypes = [C.c_void_p]
    L.sphb200_group_create_rank.argtypes = [C.POINTER(Params), C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    L.sphb200_group_destroy.argtypes = [H]
    L.sphb200_group_last_error.argtypes = [H]
    L.sphb200_group_last_error.restype = C.c_char_p
    L.sphb200_group_body_range.argtypes = [H, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.sphb200_group_upload.argtypes = [H, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.sphb200_group_step.argtypes = [H, C.c_float, C.c_int]
    L.sphb200_group_download.argtypes = [H, C.c_int, C.c_void_p, C.c_int]
    L.sphb200_group_sync.argtypes = [H]
    L.sphb200_group_diagnostics.argtypes = [H, C.c_void_p]
    L.sphb200_group_info.argtypes = [H, C.POINTER(GroupInfo)]
    L.sphb200_group_enable_timing.argtypes = [H, C.c_int]
    L.sphb200_group_get_timings.argtypes = [H, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.sphb200_group_rank_handle.argtypes = [H, C.c_int, C.POINTER(H)]
    L.sphb200_group_stream.argtypes = [H, C.c_int, C.POINTER(C.c_void_p)]
    L.sphb200_get_stream.argtypes = [H, C.POINTER(C.c_void_p)]
    L.sphb200_field_stats.argtypes = [H, C.c_void_p]
    L.sphb200_group_field_stats.argtypes = [H, C.c_void_p]
    for fn in (L.sphb200_snapshot_save, L.sphb200_snapshot_load, L.sphb200_group_snapshot_save, L.sphb200_group_snapshot_load):
        fn.argtypes = [H, C.c_char_p]
    _lib = L
    return L


def default_params(**kw):
    p = Params()
    load_library().sphb200_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Simulation:
    """One handle = one GPU-resident particle set. Thin, 1:1 over the C ABI."""

    def __init__(self, capacity, device=0, params=None, **param_overrides):
        self.L = load_library()
        self.params = params if params is not None else default_params(**param_overrides)
        self.h = C.c_void_p()
        rc = self.L.sphb200_create(C.byref(self.params), int(capacity), int(device), C.byref(self.h))
        if rc != SPH_OK:
            raise SphError(rc, (self.L.sphb200_last_error(None) or b"").decode())
        self.capacity = int(capacity)
        self.device = device
        self.n = 0

    # -- plumbing
    def _ck(self, rc, allow=()):
        if rc != SPH_OK and rc not in allow:
            raise SphError(rc, (self.L.sphb200_last_error(self.h) or b"").decode())
        return rc

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.sphb200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, stream_ptr):
        self._ck(self.L.sphb200_set_stream(self.h, C.c_void_p(stream_ptr)))

    def sync(self):
        self._ck(self.L.sphb200_sync(self.h))

    def stream_ptr(self):
        p = C.c_void_p()
        self._ck(self.L.sphb200_get_stream(self.h, C.byref(p)))
        return p.value or 0

    def effective_params(self):
        p = Params()
        self._ck(self.L.sphb200_get_params(self.h, C.byref(p)))
        return p

    # -- upload: accepts either plain arrays ((n,3) f32 pos/vel, (n,) mass/h) or the component-struct arrays
    def upload(self, pos, vel, mass, smoothing):
        pos = np.ascontiguousarray(pos); vel = np.ascontiguousarray(vel)
        mass = np.ascontiguousarray(mass); smoothing = np.ascontiguousarray(smoothing)

        def stride(a, natural):
            if a.dtype.names:
                return a.dtype.itemsize
            assert a.dtype == np.float32, "float32 arrays required"
            if len(a) == 0:
                return natural
            a2 = a.reshape(len(a), -1)
            assert a2.shape[1] * 4 >= natural
            return a2.shape[1] * 4
        n = len(mass)
        assert len(pos) == n and len(vel) == n and len(smoothing) == n
        self._ck(self.L.sphb200_upload(self.h, n, _ptr(pos), stride(pos, 12), _ptr(vel), stride(vel, 12), _ptr(mass),
                                       stride(mass, 4), _ptr(smoothing), stride(smoothing, 4)))
        self.n = n

    # -- stages (one per reference system)
    def smoothing_update(self):
        self._ck(self.L.sphb200_smoothing_update(self.h))

    def build_neighbors(self):
        self._ck(self.L.sphb200_build_neighbors(self.h))

    def gravity(self, impl, dt):
        self._ck(self.L.sphb200_gravity(self.h, int(impl), float(dt)))

    def prepare_gravity(self, impl, dt):
        self._ck(self.L.sphb200_prepare_gravity(self.h, int(impl), float(dt)))

    def density(self):
        self._ck(self.L.sphb200_density(self.h))

    def pressure(self):
        self._ck(self.L.sphb200_pressure(self.h))

    def integrate(self, dt):
        self._ck(self.L.sphb200_integrate(self.h, float(dt)))

    def step(self, dt, gravity=GRAVITY_TREE):
        self._ck(self.L.sphb200_step(self.h, float(dt), int(gravity)))

    def set_target_range(self, t0, t1):
        self._ck(self.L.sphb200_set_target_range(self.h, int(t0), int(t1)))

    # -- results
    _SHAPES = {FIELD_TRANSLATION: ("f4", 3), FIELD_VELOCITY: ("f4", 3), FIELD_MASS: ("f4", 1), FIELD_DENSITY: ("f4", 1),
               FIELD_PRESSURE: ("f4", 1), FIELD_PRESSURE_GRAD: ("f4", 3), FIELD_NEIGHBOR_COUNT: ("i4", 1)}

    def download(self, field, out=None, allow_overflow=False):
        n = self.n
        if out is None:
            if field == FIELD_SMOOTHING:
                out = np.zeros(n, ParticleSmoothing)
            elif field == FIELD_GRAVITY:
                out = np.zeros(n, GravityField)
            else:
                dt, w = self._SHAPES[field]
                out = np.zeros((n, w) if w > 1 else n, dt)
        stride = out.dtype.itemsize if out.dtype.names else out.reshape(n, -1).shape[1] * out.dtype.itemsize if n else 4
        allow = (SPH_ERR_NEIGHBOR_OVERFLOW,) if allow_overflow else ()
        self._ck(self.L.sphb200_download(self.h, int(field), _ptr(out), int(stride)), allow)
        return out

    def download_all(self):
        """dict of plain arrays: pos, vel, mass, h, n_own, rho, P, gradP, grav (n,4), num_particles, num_approx, count."""
        sm = self.download(FIELD_SMOOTHING)
        gf = self.download(FIELD_GRAVITY)
        return dict(pos=self.download(FIELD_TRANSLATION), vel=self.download(FIELD_VELOCITY), mass=self.download(FIELD_MASS),
                    h=sm["influenceArea"].copy(), n_own=sm["neighbors"].copy(), rho=self.download(FIELD_DENSITY),
                    P=self.download(FIELD_PRESSURE), gradP=self.download(FIELD_PRESSURE_GRAD), grav=gf["value"].copy(),
                    num_particles=gf["numParticles"].copy(), num_approx=gf["numApprox"].copy(),
                    count=self.download(FIELD_NEIGHBOR_COUNT))

    def download_neighbors(self, allow_overflow=False):
        n = self.n
        offsets = np.zeros(n + 1, np.int64)
        total = C.c_int64(0)
        allow = (SPH_ERR_NEIGHBOR_OVERFLOW,) if allow_overflow else ()
        self._ck(self.L.sphb200_download_neighbors(self.h, _ptr(offsets), None, 0, C.byref(total)), allow)
        nbr = np.zeros(max(total.value, 1), np.int32)
        self._ck(self.L.sphb200_download_neighbors(self.h, _ptr(offsets), _ptr(nbr), total.value, C.byref(total)), allow)
        return offsets, nbr[: total.value]

    def download_interactions(self, offsets, nbr):
        out = np.zeros(max(len(nbr), 1), ParticleInteraction)
        offsets = np.ascontiguousarray(offsets, np.int64); nbr = np.ascontiguousarray(nbr, np.int32)
        self._ck(self.L.sphb200_download_interactions(self.h, _ptr(offsets), _ptr(nbr), _ptr(out)))
        return out[: len(nbr)]

    def download_sort(self):
        order = np.zeros(self.n, np.uint32); keys = np.zeros(self.n, np.uint32); g = GridParams()
        self._ck(self.L.sphb200_download_sort(self.h, _ptr(order), _ptr(keys), C.byref(g)))
        return order, keys, g

    def download_tree(self):
        nn = 2 * self.n - 1
        child = np.zeros((nn, 2), np.int32); rng = np.zeros((nn, 2), np.int32)
        mom = np.zeros((nn, 4), np.float32); lo = np.zeros((nn, 3), np.float32); hi = np.zeros((nn, 3), np.float32)
        self._ck(self.L.sphb200_download_tree(self.h, _ptr(child), _ptr(rng), _ptr(mom), _ptr(lo), _ptr(hi)))
        return dict(child=child, range=rng, mom=mom, lo=lo, hi=hi)

    def diagnostics(self):
        out = np.zeros(12, np.float64)
        self._ck(self.L.sphb200_diagnostics(self.h, _ptr(out)))
        return dict(mass=out[0], momentum=out[1:4].copy(), angular_momentum=out[4:7].copy(), e_kin=out[7], e_pot=out[8],
                    e_int=out[9], mean_neighbors=out[10], max_neighbors=int(out[11]))

    def field_stats(self):
        """{"rho" | "P" | "grav" | "u": (min, max, mean)} over all particles (README.md:50-52 roadmap)."""
        out = np.zeros(12, np.float64)
        self._ck(self.L.sphb200_field_stats(self.h, _ptr(out)))
        return {k: tuple(out[3 * i:3 * i + 3]) for i, k in enumerate(("rho", "P", "grav", "u"))}

    # -- snapshot I/O (checkpoint/resume and the fixture format): body-order arrays in one binary file,
    #    written and read by the library (sphb200_snapshot_save / _load)
    def save_snapshot(self, path):
        self._ck(self.L.sphb200_snapshot_save(self.h, os.fsencode(path)), (SPH_ERR_NEIGHBOR_OVERFLOW,))

    def load_snapshot(self, path):
        self._ck(self.L.sphb200_snapshot_load(self.h, os.fsencode(path)))
        n = C.c_int64(0); cap = C.c_int64(0)
        self._ck(self.L.sphb200_count(self.h, C.byref(n), C.byref(cap)))
        self.n = n.value

    def launch_count(self):
        v = C.c_int64(0)
        self._ck(self.L.sphb200_launch_count(self.h, C.byref(v)))
        return v.value

    def enable_timing(self, on=True):
        self._ck(self.L.sphb200_enable_timing(self.h, 1 if on else 0))

    def timings(self):
        names = (C.c_char_p * 32)(); ms = (C.c_float * 32)()
        k = self.L.sphb200_get_timings(self.h, names, ms, 32)
        return [(names[i].decode(), float(ms[i])) for i in range(max(k, 0))]

    def fp32_peak_tflops(self):
        v = C.c_double(0)
        self._ck(self.L.sphb200_fp32_peak(self.h, C.byref(v)))
        return v.value

    def device_ptr(self, name):
        p = C.c_void_p(); b = C.c_int64(0)
        self._ck(self.L.sphb200_device_ptr(self.h, name.encode(), C.byref(p), C.byref(b)))
        return p.value, b.value
