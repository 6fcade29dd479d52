"""Seeded synthetic initial conditions (replaces the reference's unseeded spawner).

Reference: Assets/Scripts/Systems/ParticleAuthoring.cs:150-245 -- positions uniform in a ball by rejection
from the cube (:229-245), equal masses M/N (:208), zero velocity (:165), support radius
``particleRadius*(1+U[0,0.5))`` hence ``h0 = particleRadius*(1+U[0,0.5))/2`` (:164, ParticleSmoothing.cs:9-15).
The reference seeds from an unseeded System.Random (RandomSystem.cs:36-40, quirk Q11) so its ICs are not
reproducible; here a counter-based Philox stream makes every configuration bit-reproducible.

Scene values of the reference (Assets/Scenes/SimScene.unity:276-279): count 3000, particleRadius 5,
radius 50, totalMass 100.  ``particle_radius=None`` scales 5.0 with R*N^(-1/3) so that number of neighbors
per support volume stays that of the reference scene at any N.
"""
import numpy as np

REF_COUNT, REF_PARTICLE_RADIUS, REF_RADIUS, REF_TOTAL_MASS = 3000, 5.0, 50.0, 100.0


def _ball(rng, n, radius):
    out = np.empty((n, 3), np.float32)
    filled = 0
    while filled < n:
        m = int((n - filled) * 2.0) + 16
        c = rng.uniform(-radius, radius, size=(m, 3)).astype(np.float32)
        ok = (c.astype(np.float64) ** 2).sum(1) <= float(radius) ** 2
        c = c[ok][: n - filled]
        out[filled:filled + len(c)] = c
        filled += len(c)
    return out


def particle_radius_for(n, radius):
    """Support radius that reproduces the reference scene's neighbor loading at any (N, R)."""
    k = REF_PARTICLE_RADIUS / (REF_RADIUS * REF_COUNT ** (-1.0 / 3.0))  # = 1.442
    return float(k * radius * n ** (-1.0 / 3.0))


def make_sphere(n, radius=REF_RADIUS, total_mass=REF_TOTAL_MASS, particle_radius=None, seed=1234,
                center=(0.0, 0.0, 0.0), velocity=(0.0, 0.0, 0.0)):
    """Returns dict(pos (n,3) f32, vel (n,3) f32, mass (n,) f32, h (n,) f32)."""
    rng = np.random.Generator(np.random.Philox(seed))
    if particle_radius is None:
        particle_radius = particle_radius_for(n, radius)
    pos = _ball(rng, n, radius) + np.asarray(center, np.float32)
    vel = np.tile(np.asarray(velocity, np.float32), (n, 1))
    inst = (particle_radius * (1.0 + rng.uniform(0.0, 0.5, size=n))).astype(np.float32)
    h = (inst / np.float32(2.0)).astype(np.float32)
    mass = np.full(n, np.float32(total_mass) / np.float32(n), np.float32)
    return dict(pos=np.ascontiguousarray(pos, np.float32), vel=np.ascontiguousarray(vel), mass=mass, h=h)


def scaled_radius(n):
    """Sphere radius that keeps the reference scene's number density when N grows (R ~ N^(1/3))."""
    return float(REF_RADIUS * (n / REF_COUNT) ** (1.0 / 3.0))


def make_config(name, seed=1234, particles=None):
    """Named BASELINE.json configurations (`particles` overrides the size of c3/c4/c5 for scaled-down runs)."""
    if particles is not None and name in ("c3", "c4"):
        n = int(particles)
        return make_sphere(n, radius=scaled_radius(n), total_mass=REF_TOTAL_MASS * n / REF_COUNT, seed=seed)
    if name == "c5":      # two-planet collision, 2 x 4M, 64x density contrast
        return make_collision(int(particles) // 2 if particles else 4_000_000, seed=seed)
    if name == "c1":      # Jupiter v1: 3k, direct gravity
        return make_sphere(3000, seed=seed)
    if name == "c2":      # 10k, tree gravity
        return make_sphere(10000, seed=seed)
    if name == "c3":      # 1M uniform sphere (density of the reference scene, R scaled)
        n = 1 << 20
        return make_sphere(n, radius=scaled_radius(n), total_mass=REF_TOTAL_MASS * n / REF_COUNT, seed=seed)
    if name == "c4":      # 16M
        n = 16_000_000
        return make_sphere(n, radius=scaled_radius(n), total_mass=REF_TOTAL_MASS * n / REF_COUNT, seed=seed)
    raise ValueError(name)


def make_rotating_sphere(n, omega=0.02, seed=1234):
    """Rigidly over-rotating sphere about z (README.md:67-71 roadmap scenario)."""
    c = make_sphere(n, radius=scaled_radius(n), total_mass=REF_TOTAL_MASS * n / REF_COUNT, seed=seed)
    c["vel"] = np.ascontiguousarray(np.stack([-omega * c["pos"][:, 1], omega * c["pos"][:, 0], np.zeros(n)], 1).astype(np.float32))
    return c


def make_collision(n_each, seed=1234, separation=3.0, v0=1.0):
    """Two-planet collision (config C5): second body has 8x the mass in 0.5x the radius (64x density)."""
    r1 = scaled_radius(n_each)
    a = make_sphere(n_each, radius=r1, total_mass=REF_TOTAL_MASS * n_each / REF_COUNT, seed=seed,
                    center=(-0.5 * separation * r1, 0, 0), velocity=(v0, 0, 0))
    b = make_sphere(n_each, radius=0.5 * r1, total_mass=8 * REF_TOTAL_MASS * n_each / REF_COUNT, seed=seed + 1,
                    center=(0.5 * separation * r1, 0, 0), velocity=(-v0, 0, 0))
    return {k: np.ascontiguousarray(np.concatenate([a[k], b[k]])) for k in a}


def polytrope_radius(K=1000.0, G=1.0):
    """Equilibrium radius of the n = 1 polytrope that the reference's EOS P = K rho^2 (PressureFieldSystem.cs:31-33)
    supports against self-gravity: R = pi sqrt(K / (2 pi G)), independent of the mass (39.63 for K = 1000, G = 1)."""
    return float(np.pi * np.sqrt(K / (2.0 * np.pi * G)))


def make_polytrope(n, total_mass=REF_TOTAL_MASS, K=1000.0, G=1.0, neighbors=50.0, seed=1234, center=(0.0, 0.0, 0.0),
                   velocity=(0.0, 0.0, 0.0), omega=0.0):
    """Non-uniform density initial condition (README.md:85-89 roadmap: "nonuniform density", "automatic particle size",
    "initial group velocity / angular momentum"): equal-mass particles drawn from the hydrostatic n = 1 polytrope
    rho(r) = rho_c sin(xi)/xi, xi = pi r / R, R = polytrope_radius(K, G) -- the equilibrium of the reference's own EOS, so
    the sphere starts near rest instead of ringing like the uniform ball.  Radii by inverse transform of the enclosed mass
    m(r)/M = (sin(xi) - xi cos(xi)) / pi, directions isotropic.  h is set from the local density so that ~`neighbors`
    particles lie inside 2h (the controller's target, ParticleSmoothingSystem.cs:18).  `omega` adds rigid rotation about z."""
    rng = np.random.Generator(np.random.Philox(seed))
    R = polytrope_radius(K, G)
    # inverse transform on a fine table of the (monotone) enclosed-mass fraction
    xi_t = np.linspace(0.0, np.pi, 20001)
    m_t = (np.sin(xi_t) - xi_t * np.cos(xi_t)) / np.pi
    xi = np.interp(rng.uniform(0.0, 1.0, n), m_t, xi_t)
    r = xi * (R / np.pi)
    mu = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    st = np.sqrt(1.0 - mu * mu)
    pos = np.stack([r * st * np.cos(ph), r * st * np.sin(ph), r * mu], 1)
    rho_c = total_mass * np.pi / (4.0 * R ** 3)                     # M = 4 R^3 rho_c / pi
    with np.errstate(invalid="ignore", divide="ignore"):
        rho = rho_c * np.where(xi > 1e-6, np.sin(xi) / xi, 1.0)
    rho = np.maximum(rho, rho_c * 1e-3)                             # the surface density -> 0: cap the smoothing length
    m = total_mass / n
    h = 0.5 * (3.0 * neighbors * m / (4.0 * np.pi * rho)) ** (1.0 / 3.0)
    vel = np.tile(np.asarray(velocity, np.float64), (n, 1))
    if omega:
        vel = vel + np.stack([-omega * pos[:, 1], omega * pos[:, 0], np.zeros(n)], 1)
    pos = pos + np.asarray(center, np.float64)
    return dict(pos=np.ascontiguousarray(pos, np.float32), vel=np.ascontiguousarray(vel, np.float32),
                mass=np.full(n, np.float32(m), np.float32), h=np.ascontiguousarray(h, np.float32))
