"""Multi-GPU groups over the C ABI (``sphb200_group_*``): Morton-range domain decomposition with NVLink halo exchange.

Two ways to form a group (include/sphb200.h):

* ``Group.single_process(n_capacity, devices)`` -- one process drives every GPU (what the C# host of INTEGRATION.md
  does).  Listing a device more than once puts several ranks on one GPU and switches the transport from NCCL to
  in-process peer copies: that is how the decomposition is tested on a single-GPU box.
* ``Group.from_env(n_capacity)`` -- one process per GPU under ``torchrun``: rank 0 creates the NCCL unique id, the
  128 bytes travel through ``torch.distributed`` (any backend), every process joins with its rank.

Every process uploads / downloads only its body-order slice ``[body0, body0 + count)``.
"""
import ctypes as C
import os
import numpy as np

from . import (FIELD_DENSITY, FIELD_GRAVITY, FIELD_MASS, FIELD_NEIGHBOR_COUNT, FIELD_PRESSURE, FIELD_PRESSURE_GRAD,
               FIELD_SMOOTHING, FIELD_TRANSLATION, FIELD_VELOCITY, GRAVITY_TREE, GravityField, GroupInfo, ParticleSmoothing,
               SPH_ERR_NEIGHBOR_OVERFLOW, SPH_OK, SphError, default_params, load_library, _ptr)


class Group:
    def __init__(self, handle, n_capacity, params):
        self.L = load_library()
        self.h = handle
        self.capacity = int(n_capacity)
        self.params = params
        self.n_total = 0
        self.body0 = 0
        self.count = 0

    # ---- construction
    @classmethod
    def single_process(cls, n_capacity, devices, params=None, **param_overrides):
        L = load_library()
        p = params if params is not None else default_params(**param_overrides)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = L.sphb200_group_create(C.byref(p), int(n_capacity), len(devices), devs, C.byref(h))
        if rc != SPH_OK:
            raise SphError(rc, (L.sphb200_group_last_error(None) or b"").decode())
        return cls(h, n_capacity, p)

    @classmethod
    def from_env(cls, n_capacity, device=None, params=None, **param_overrides):
        """One rank per process (RANK / WORLD_SIZE / LOCAL_RANK from torchrun); torch.distributed must be initialised when
        WORLD_SIZE > 1 (it only carries the 128-byte NCCL unique id)."""
        L = load_library()
        p = params if params is not None else default_params(**param_overrides)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        uid = (C.c_ubyte * 128)()
        if world > 1:
            import torch
            import torch.distributed as td
            if rank == 0:
                rc = L.sphb200_group_unique_id(uid)
                if rc != SPH_OK:
                    raise SphError(rc, (L.sphb200_group_last_error(None) or b"").decode())
            t = torch.tensor(list(uid), dtype=torch.uint8)
            if td.get_backend() == "nccl":
                t = t.cuda(device)
            td.broadcast(t, 0)
            uid = (C.c_ubyte * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        rc = L.sphb200_group_create_rank(C.byref(p), int(n_capacity), uid, world, rank, int(device), C.byref(h))
        if rc != SPH_OK:
            raise SphError(rc, (L.sphb200_group_last_error(None) or b"").decode())
        return cls(h, n_capacity, p)

    # ---- plumbing
    def _ck(self, rc, allow=()):
        if rc != SPH_OK and rc not in allow:
            raise SphError(rc, (self.L.sphb200_group_last_error(self.h) or b"").decode())
        return rc

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.sphb200_group_destroy(self.h)
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

    def body_range(self, n_total):
        b0 = C.c_int64(0); cnt = C.c_int64(0)
        self._ck(self.L.sphb200_group_body_range(self.h, int(n_total), C.byref(b0), C.byref(cnt)))
        return b0.value, cnt.value

    # ---- data
    def upload(self, n_total, pos, vel, mass, smoothing):
        """Arrays of THIS process's bodies [body0, body0+count) (see body_range); plain float32 arrays or component structs."""
        pos = np.ascontiguousarray(pos); vel = np.ascontiguousarray(vel)
        mass = np.ascontiguousarray(mass); smoothing = np.ascontiguousarray(smoothing)

        def stride(a, natural):
            if a.dtype.names:
                return a.dtype.itemsize
            assert a.dtype == np.float32, "float32 arrays required"
            return natural if len(a) == 0 else a.reshape(len(a), -1).shape[1] * 4
        self.body0, self.count = self.body_range(n_total)
        assert len(pos) == self.count == len(vel) == len(mass) == len(smoothing), "pass this process's body slice"
        self._ck(self.L.sphb200_group_upload(self.h, int(n_total), _ptr(pos), stride(pos, 12), _ptr(vel), stride(vel, 12), _ptr(mass),
                                             stride(mass, 4), _ptr(smoothing), stride(smoothing, 4)))
        self.n_total = int(n_total)

    def upload_global(self, pos, vel, mass, smoothing):
        """Convenience: every process holds the full arrays and uploads its slice of them."""
        n = len(mass)
        b0, cnt = self.body_range(n)
        self.upload(n, pos[b0:b0 + cnt], vel[b0:b0 + cnt], mass[b0:b0 + cnt], smoothing[b0:b0 + cnt])

    def step(self, dt, gravity=GRAVITY_TREE):
        self._ck(self.L.sphb200_group_step(self.h, float(dt), int(gravity)))

    def sync(self):
        self._ck(self.L.sphb200_group_sync(self.h))

    def stream_ptr(self, local_rank=0):
        p = C.c_void_p()
        self._ck(self.L.sphb200_group_stream(self.h, int(local_rank), C.byref(p)))
        return p.value or 0

    _SHAPES = {FIELD_TRANSLATION: ("f4", 3), FIELD_VELOCITY: ("f4", 3), FIELD_MASS: ("f4", 1), FIELD_DENSITY: ("f4", 1),
               FIELD_PRESSURE: ("f4", 1), FIELD_PRESSURE_GRAD: ("f4", 3), FIELD_NEIGHBOR_COUNT: ("i4", 1)}

    def download(self, field, out=None, allow_overflow=False):
        n = self.count
        if out is None:
            if field == FIELD_SMOOTHING:
                out = np.zeros(n, ParticleSmoothing)
            elif field == FIELD_GRAVITY:
                out = np.zeros(n, GravityField)
            else:
                dt, w = self._SHAPES[field]
                out = np.zeros((n, w) if w > 1 else n, dt)
        stride = out.dtype.itemsize if out.dtype.names else (out.reshape(n, -1).shape[1] * out.dtype.itemsize if n else 4)
        allow = (SPH_ERR_NEIGHBOR_OVERFLOW,) if allow_overflow else ()
        self._ck(self.L.sphb200_group_download(self.h, int(field), _ptr(out), int(stride)), allow)
        return out

    def download_all(self):
        """Same dict as Simulation.download_all, for this process's body slice."""
        sm = self.download(FIELD_SMOOTHING)
        gf = self.download(FIELD_GRAVITY)
        return dict(pos=self.download(FIELD_TRANSLATION), vel=self.download(FIELD_VELOCITY), mass=self.download(FIELD_MASS),
                    h=sm["influenceArea"].copy(), n_own=sm["neighbors"].copy(), rho=self.download(FIELD_DENSITY),
                    P=self.download(FIELD_PRESSURE), gradP=self.download(FIELD_PRESSURE_GRAD), grav=gf["value"].copy(),
                    num_particles=gf["numParticles"].copy(), num_approx=gf["numApprox"].copy(),
                    count=self.download(FIELD_NEIGHBOR_COUNT))

    def diagnostics(self):
        out = np.zeros(12, np.float64)
        self._ck(self.L.sphb200_group_diagnostics(self.h, _ptr(out)))
        return dict(mass=out[0], momentum=out[1:4].copy(), angular_momentum=out[4:7].copy(), e_kin=out[7], e_pot=out[8],
                    e_int=out[9], mean_neighbors=out[10], max_neighbors=int(out[11]))

    def field_stats(self):
        out = np.zeros(12, np.float64)
        self._ck(self.L.sphb200_group_field_stats(self.h, _ptr(out)))
        return {k: tuple(out[3 * i:3 * i + 3]) for i, k in enumerate(("rho", "P", "grav", "u"))}

    def save_snapshot(self, path):
        self._ck(self.L.sphb200_group_snapshot_save(self.h, os.fsencode(path)), (SPH_ERR_NEIGHBOR_OVERFLOW,))

    def load_snapshot(self, path):
        self._ck(self.L.sphb200_group_snapshot_load(self.h, os.fsencode(path)))
        gi = GroupInfo()
        self._ck(self.L.sphb200_group_info(self.h, C.byref(gi)))
        self.n_total = int(gi.n_total)
        self.body0, self.count = self.body_range(self.n_total)

    def info(self):
        gi = GroupInfo()
        self._ck(self.L.sphb200_group_info(self.h, C.byref(gi)))
        return dict(world=gi.world, nlocal=gi.nlocal, rank0=gi.rank0, transport=("nccl", "local", "none")[gi.transport],
                    n_total=gi.n_total, steps=gi.steps, migrated_last_step=gi.migrated_last_step, halo_last_step=gi.halo_last_step,
                    cap_own=gi.cap_own, cap_halo=gi.cap_halo, launches=gi.launches, tree_nodes_last_step=gi.tree_nodes_last_step, n_own=list(gi.n_own[:gi.nlocal]),
                    n_halo=list(gi.n_halo[:gi.nlocal]))

    def launch_count(self):
        return self.info()["launches"]

    def enable_timing(self, on=True):
        self._ck(self.L.sphb200_group_enable_timing(self.h, 1 if on else 0))

    def timings(self):
        names = (C.c_char_p * 32)(); ms = (C.c_float * 32)()
        k = self.L.sphb200_group_get_timings(self.h, names, ms, 32)
        return [(names[i].decode(), float(ms[i])) for i in range(max(k, 0))]
