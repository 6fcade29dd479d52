"""ctypes front-end of the CPU oracle (oracle/sph_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (planetmodel-sph_b200/) never imports this module.

All arrays are numpy float32/int32, positions/velocities are (N,3) C-contiguous.
The step composition follows the reference's system order:
  ParticleSmoothingSystem -> KernelSystem (+GravityFieldSystem) -> Integrator (x += v dt)
  -> DensityFieldSystem -> PressureFieldSystem -> VelocitySystem (v += a dt).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


class GridParams(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("cell", C.c_float), ("fine_scale", C.c_float), ("bits", C.c_int32),
                ("hmax", C.c_float), ("ext", C.c_float), ("href", C.c_float), ("stencil", C.c_int32)]


def build(force=False):
    """Compile liborc.so in place (g++ only; no reference sources are compiled -- the reference is C#)."""
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "sph_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_kernel.restype = C.c_float
        L.orc_kernel.argtypes = [C.c_float, C.c_float]
        L.orc_kernel_deriv.restype = C.c_float
        L.orc_kernel_deriv.argtypes = [C.c_float, C.c_float, C.c_int]
        L.orc_interacts.argtypes = [f32p, f32p, C.c_float, C.c_float]
        L.orc_is_neighbor.argtypes = [f32p, f32p, C.c_float, C.c_float]
        L.orc_kernel_and_gradient.argtypes = [f32p, f32p, C.c_float, C.c_int, f32p]
        L.orc_interaction.argtypes = [f32p, f32p, C.c_float, C.c_float, C.c_int, f32p]
        L.orc_gravity_pair.argtypes = [f32p, f32p, C.c_float, C.c_float, C.c_float, f32p]
        L.orc_moment_accumulate.argtypes = [f32p, f32p, C.c_float]
        L.orc_moment_m2p.argtypes = [f32p, f32p, C.c_float, f32p]
        L.orc_accept.argtypes = [f32p, f32p, f32p, f32p, C.c_float]
        L.orc_expand_aabb.argtypes = [f32p, f32p, f32p, C.c_float, f32p, f32p]
        L.orc_calculate_expansion.argtypes = [f32p, f32p, C.c_float, C.c_float, f32p, f32p]
        L.orc_particle_box.argtypes = [f32p, C.c_float, f32p, C.c_float, C.c_int, f32p, f32p]
        L.orc_radius_ratio.restype = C.c_float
        L.orc_radius_ratio.argtypes = [C.c_float, C.c_int]
        L.orc_smoothing_update.argtypes = [C.c_int64, f32p, i32p, C.c_float, f32p]
        for fn in (L.orc_neighbors_brute, L.orc_neighbors_grid):
            fn.restype = C.c_int64
            fn.argtypes = [C.c_int64, f32p, f32p, i64p, i32p, C.c_int64]
        L.orc_interactions.argtypes = [C.c_int64, f32p, f32p, i64p, i32p, C.c_int, f32p, f32p]
        L.orc_density.argtypes = [C.c_int64, f32p, f32p, f32p, i64p, i32p, f32p, i32p]
        L.orc_eos.argtypes = [C.c_int64, f32p, C.c_float, f32p]
        L.orc_pressure_grad.argtypes = [C.c_int64, f32p, f32p, f32p, f32p, f32p, i64p, i32p, C.c_int, f32p]
        L.orc_gravity_direct.argtypes = [C.c_int64, f32p, f32p, f32p, C.c_float, C.c_int64, C.c_int64, C.c_int, f32p]
        L.orc_gravity_direct_pm07.argtypes = [C.c_int64, f32p, f32p, f32p, C.c_float, C.c_int64, C.c_int64, f32p]
        L.orc_gravity_pm07_correction.argtypes = [C.c_int64, f32p, f32p, f32p, C.c_float, i64p, i32p, f32p]
        L.orc_pm07_kernel.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.orc_integrate.argtypes = [C.c_int64, f32p, f32p, f32p, f32p, f32p, C.c_float]
        L.orc_integrate2.argtypes = [C.c_int64, f32p, f32p, f32p, f32p, f32p, C.c_float, C.c_int]
        L.orc_grid_params.argtypes = [C.c_int64, f32p, f32p, C.c_int, C.POINTER(GridParams)]
        L.orc_morton_keys.argtypes = [C.c_int64, f32p, C.POINTER(GridParams), u32p]
        L.orc_sort_order.argtypes = [C.c_int64, u32p, u32p]
        L.orc_lbvh_topology.argtypes = [C.c_int64, u32p, i32p, i32p, i32p, i32p, i32p]
        L.orc_lbvh_moments.argtypes = [C.c_int64, f32p, f32p, f32p, f32p, i32p, i32p, i32p, i32p, C.c_int, C.c_int,
                                       C.c_float, f32p, f32p, f32p]
        L.orc_tree_walk.argtypes = [C.c_int64, f32p, f32p, f32p, i32p, i32p, i32p, i32p, f32p, f32p, f32p, C.c_int,
                                    C.c_float, C.c_float, C.c_int64, C.c_int64, C.c_int, f32p, i32p, i32p]
        L.orc_tree4_gravity.argtypes = [C.c_int64, f32p, f32p, f32p, f32p, C.c_float, C.c_float, C.c_float, C.c_int, f32p, i32p, i32p,
                                        C.POINTER(C.c_int32)]
        L.orc_reference_step.restype = C.c_int64
        L.orc_reference_step.argtypes = [C.c_int64, f32p, f32p, f32p, f32p, i32p, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float,
                                         C.c_float, C.c_int, f32p, f32p, f32p, f32p, i32p, i32p, i64p, C.c_void_p, C.c_int64, i64p,
                                         np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------- scalar helpers
def kernel(r, h):
    return float(lib().orc_kernel(r, h))


def kernel_deriv(r, h, fix_q1=0):
    return float(lib().orc_kernel_deriv(r, h, fix_q1))


def interaction(ri, rj, hi, hj, fix_q1=0):
    out = np.zeros(8, np.float32)
    lib().orc_interaction(_f(ri), _f(rj), hi, hj, fix_q1, out)
    return out[:4].copy(), out[4:].copy()


def gravity_pair(ri, rj, m, a, G=1.0):
    out = np.zeros(4, np.float32)
    lib().orc_gravity_pair(_f(ri), _f(rj), m, a, G, out)
    return out


# ---------------------------------------------------------------- neighbor sets
def neighbors(pos, h, method="auto"):
    """CSR neighbor lists (offsets int64[n+1], nbr int32[total]), ascending j per particle."""
    pos = _f(pos); h = _f(h); n = len(h)
    L = lib()
    fn = L.orc_neighbors_brute if (method == "brute" or (method == "auto" and n <= 4096)) else L.orc_neighbors_grid
    offsets = np.zeros(n + 1, np.int64)
    cap = max(64 * n, 1)
    while True:
        nbr = np.zeros(cap, np.int32)
        tot = fn(n, pos, h, offsets, nbr, cap)
        if tot <= cap:
            return offsets, nbr[:tot].copy()
        cap = int(tot)


def density(pos, h, m, offsets, nbr):
    pos = _f(pos); h = _f(h); m = _f(m); n = len(h)
    rho = np.zeros(n, np.float32); own = np.zeros(n, np.int32)
    lib().orc_density(n, pos, h, m, offsets, nbr, rho, own)
    return rho, own


def eos(rho, K=1000.0):
    P = np.zeros_like(rho)
    lib().orc_eos(len(rho), _f(rho), K, P)
    return P


def pressure_grad(pos, h, m, rho, P, offsets, nbr, fix_q1=0):
    n = len(h); g = np.zeros((n, 3), np.float32)
    lib().orc_pressure_grad(n, _f(pos), _f(h), _f(m), _f(rho), _f(P), offsets, nbr, fix_q1, g)
    return g


def gravity_direct(pos, h, m, G=1.0, i0=0, i1=None, accum_double=False):
    n = len(h); i1 = n if i1 is None else i1
    g = np.zeros((i1 - i0, 4), np.float32)
    lib().orc_gravity_direct(n, _f(pos), _f(h), _f(m), G, i0, i1, int(accum_double), g)
    return g


def gravity_direct_pm07(pos, h, m, G=1.0, i0=0, i1=None):
    """NON-REFERENCE option (README.md:75-77 roadmap): Price & Monaghan 2007 spline-softened gravity, symmetric in h_i, h_j."""
    n = len(h); i1 = n if i1 is None else i1
    g = np.zeros((i1 - i0, 4), np.float32)
    lib().orc_gravity_direct_pm07(n, _f(pos), _f(h), _f(m), G, i0, i1, g)
    return g


def gravity_pm07_correction(pos, h, m, offsets, nbr, G=1.0):
    """Sum over the neighbor lists of (PM07 pair law - reference pair law): what SPH_FLAG_PM07_SOFTENING adds to a tree walk."""
    n = len(h); g = np.zeros((n, 4), np.float32)
    lib().orc_gravity_pm07_correction(n, _f(pos), _f(h), _f(m), G, offsets, nbr, g)
    return g


def pm07_kernel(r, h):
    out = (C.c_double * 2)()
    lib().orc_pm07_kernel(r, h, out)
    return out[0], out[1]


def smoothing_update(h, n_own, target=50.0):
    out = np.zeros_like(_f(h))
    lib().orc_smoothing_update(len(out), _f(h), np.ascontiguousarray(n_own, np.int32), target, out)
    return out


def integrate(pos, vel, rho, gradP, grav4, dt, kick_drift=False):
    pos = _f(pos).copy(); vel = _f(vel).copy()
    lib().orc_integrate2(len(rho), pos, vel, _f(rho), _f(gradP), _f(grav4), dt, int(kick_drift))
    return pos, vel


# ---------------------------------------------------------------- keys / sort / LBVH
def grid_params(pos, h, max_bits):
    g = GridParams()
    lib().orc_grid_params(len(h), _f(pos), _f(h), max_bits, C.byref(g))
    return g


def morton_keys(pos, g):
    keys = np.zeros(len(pos), np.uint32)
    lib().orc_morton_keys(len(pos), _f(pos), C.byref(g), keys)
    return keys


def sort_order(keys):
    order = np.zeros(len(keys), np.uint32)
    lib().orc_sort_order(len(keys), np.ascontiguousarray(keys, np.uint32), order)
    return order


class Lbvh:
    pass


def lbvh_build(keys_sorted, pos_s, vel_s, h_s, m_s, leaf_max=4, aabb_mode=0, dt=0.0):
    n = len(keys_sorted); nn = 2 * n - 1
    t = Lbvh()
    t.n = n
    t.left = np.full(nn, -1, np.int32); t.right = np.full(nn, -1, np.int32); t.parent = np.full(nn, -1, np.int32)
    t.first = np.zeros(nn, np.int32); t.last = np.zeros(nn, np.int32)
    lib().orc_lbvh_topology(n, np.ascontiguousarray(keys_sorted, np.uint32), t.left, t.right, t.parent, t.first, t.last)
    t.mom = np.zeros((nn, 4), np.float32); t.lo = np.zeros((nn, 3), np.float32); t.hi = np.zeros((nn, 3), np.float32)
    lib().orc_lbvh_moments(n, _f(pos_s), _f(vel_s), _f(h_s), _f(m_s), t.left, t.right, t.first, t.last, leaf_max,
                           aabb_mode, dt, t.mom, t.lo, t.hi)
    t.leaf_max = leaf_max
    return t


def tree_walk(tree, pos_s, h_s, m_s, theta=0.7, G=1.0, t0=0, t1=None, accum_double=False):
    n = tree.n; t1 = n if t1 is None else t1
    g = np.zeros((t1 - t0, 4), np.float32); npart = np.zeros(t1 - t0, np.int32); napp = np.zeros(t1 - t0, np.int32)
    lib().orc_tree_walk(n, _f(pos_s), _f(h_s), _f(m_s), tree.left, tree.right, tree.first, tree.last, tree.mom,
                        tree.lo, tree.hi, tree.leaf_max, theta, G, t0, t1, int(accum_double), g, npart, napp)
    return g, npart, napp


def tree_gravity(pos, vel, h, m, dt, theta=0.7, G=1.0, leaf_max=4, aabb_mode=0, max_bits=7, accum_double=False):
    """Tree gravity for particles in ORIGINAL order: keys -> stable sort -> LBVH -> walk -> unsort."""
    pos = _f(pos); vel = _f(vel); h = _f(h); m = _f(m)
    g = grid_params(pos, h, max_bits)
    keys = morton_keys(pos, g)
    order = sort_order(keys).astype(np.int64)
    tree = lbvh_build(keys[order], pos[order], vel[order], h[order], m[order], leaf_max, aabb_mode, dt)
    gs, npart, napp = tree_walk(tree, pos[order], h[order], m[order], theta, G, accum_double=accum_double)
    out = np.zeros_like(gs); out[order] = gs
    onp = np.zeros_like(npart); onp[order] = npart
    ona = np.zeros_like(napp); ona[order] = napp
    return out, onp, ona, tree, order


def tree4_gravity(pos, vel, h, m, dt, theta=0.7, G=1.0, accum_double=False):
    """Tree gravity on the reference-SHAPED 4-ary BVH (accuracy-envelope yardstick)."""
    n = len(h)
    g = np.zeros((n, 4), np.float32); npart = np.zeros(n, np.int32); napp = np.zeros(n, np.int32); nn = C.c_int32(0)
    lib().orc_tree4_gravity(n, _f(pos), _f(vel), _f(h), _f(m), dt, theta, G, int(accum_double), g, npart, napp, C.byref(nn))
    return g, npart, napp, nn.value


# ---------------------------------------------------------------- full step
class State:
    """Per-particle state in the caller's (body-index) order."""

    def __init__(self, pos, vel, mass, h, n_own=None):
        self.pos = _f(pos).copy(); self.vel = _f(vel).copy(); self.mass = _f(mass).copy(); self.h = _f(h).copy()
        n = len(self.h)
        self.n_own = np.zeros(n, np.int32) if n_own is None else np.ascontiguousarray(n_own, np.int32).copy()
        self.rho = np.zeros(n, np.float32); self.P = np.zeros(n, np.float32)
        self.gradP = np.zeros((n, 3), np.float32); self.grav = np.zeros((n, 4), np.float32)
        self.num_particles = np.zeros(n, np.int32); self.num_approx = np.zeros(n, np.int32)
        self.offsets = None; self.nbr = None

    def copy(self):
        s = State(self.pos, self.vel, self.mass, self.h, self.n_own)
        return s


def step(s, dt, gravity="direct", K=1000.0, G=1.0, theta=0.7, target=50.0, leaf_max=4, aabb_mode=0, max_bits=7,
         accum_double=False, neighbor_method="auto", kick_drift=False):
    """One reference timestep, in place. gravity in {"direct","tree","none"}."""
    # 1. ParticleSmoothingSystem: h from last step's own-support counts (quirk Q8)
    s.h = smoothing_update(s.h, s.n_own, target)
    # 3. KernelSystem: neighbor sets at x_n, h_n
    s.offsets, s.nbr = neighbors(s.pos, s.h, neighbor_method)
    # 4. GravityFieldSystem at x_n
    if gravity == "direct":
        s.grav = gravity_direct(s.pos, s.h, s.mass, G, accum_double=accum_double)
        s.num_particles[:] = 0; s.num_approx[:] = 0
    elif gravity == "tree":
        s.grav, s.num_particles, s.num_approx, _, _ = tree_gravity(s.pos, s.vel, s.h, s.mass, dt, theta, G, leaf_max,
                                                                   aabb_mode, max_bits, accum_double)
    else:
        s.grav = np.zeros((len(s.h), 4), np.float32)
    # 6./7. Density, EOS, pressure gradient from the interactions evaluated at x_n
    s.rho, s.n_own = density(s.pos, s.h, s.mass, s.offsets, s.nbr)
    s.P = eos(s.rho, K)
    s.gradP = pressure_grad(s.pos, s.h, s.mass, s.rho, s.P, s.offsets, s.nbr)
    # 5f + 9. x += v_n dt ; v += a dt
    s.pos, s.vel = integrate(s.pos, s.vel, s.rho, s.gradP, s.grav, dt, kick_drift)
    return s


# ---------------------------------------------------------------- the reference JOB PATH (CPU baseline, bench.py)
REFERENCE_STAGES = ("smoothing", "aabb_bvh_build", "tree_overlap_candidates", "filter_pairs", "flatten_pairs", "counting_sorts_x2",
                    "calculate_interactions", "gravity", "integrate_position", "density", "eos_pressure_gradient", "velocity_cleanup")


def reference_step(s, dt, gravity="direct", K=1000.0, G=1.0, theta=0.7, target=50.0, fix_q1=0, want_lists=True):
    """One step of the reference's job path with its own structure (orc_reference_step): candidate pairs from the dual-tree
    self-overlap of the Unity-shaped 4-ary BVH, FilterPairs, flatten, the two single-thread counting sorts, the interaction
    buffers with both kernels evaluated on either side, then gravity / density / pressure / integration.  Updates the State
    in place like step(); returns {"stage_sec": {...}, "candidates", "pairs", "interactions"}.  Lists come back in the
    reference's emission order (not ascending)."""
    n = len(s.h)
    offsets = np.zeros(n + 1, np.int64)
    counts = np.zeros(3, np.int64); sec = np.zeros(12, np.float64)
    cap = max(96 * n, 1) if want_lists else 0
    nbr = np.zeros(cap, np.int32) if want_lists else None
    gcode = {"none": 0, "direct": 1, "tree": 2}[gravity]
    s.rho = np.zeros(n, np.float32); s.P = np.zeros(n, np.float32); s.gradP = np.zeros((n, 3), np.float32)
    s.grav = np.zeros((n, 4), np.float32)
    tot = lib().orc_reference_step(n, s.pos, s.vel, s.mass, s.h, s.n_own, dt, gcode, K, G, theta, target, fix_q1, s.rho, s.P, s.gradP,
                                   s.grav, s.num_particles, s.num_approx, offsets, nbr.ctypes.data if want_lists else None, cap, counts, sec)
    s.offsets = offsets
    s.nbr = nbr[:tot].copy() if (want_lists and tot <= cap) else None
    return {"stage_sec": dict(zip(REFERENCE_STAGES, sec.tolist())), "candidates": int(counts[0]), "pairs": int(counts[1]),
            "interactions": int(counts[2])}
