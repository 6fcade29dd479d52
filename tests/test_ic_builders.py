"""CPU tests of the scenario / initial-condition builders (reference spawner
Assets/Scripts/Systems/ParticleAuthoring.cs:150-245 and the README.md:67-71, 85-89 roadmap scenarios)."""
import numpy as np


def test_sphere_is_seeded_and_matches_the_reference_scene_loading():
    from sphb200 import ic
    a, b = ic.make_sphere(3000, seed=7), ic.make_sphere(3000, seed=7)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])                     # bit-reproducible (the reference's spawner is not, Q11)
    assert not np.array_equal(a["pos"], ic.make_sphere(3000, seed=8)["pos"])
    r = np.linalg.norm(a["pos"], axis=1)
    assert r.max() <= 50.0 and abs((r ** 3).mean() / 50.0 ** 3 - 0.5) < 0.03   # uniform in the ball: <r^3> = R^3/2
    assert np.all(a["mass"] == np.float32(100.0) / np.float32(3000))
    assert a["h"].min() >= 2.5 and a["h"].max() < 3.75                # particleRadius 5 * (1 + U[0,0.5)) / 2
    assert np.all(a["vel"] == 0)


def test_collision_has_the_density_contrast_and_closing_velocity():
    from sphb200 import ic
    c = ic.make_collision(5000, seed=3, separation=3.0, v0=1.0)
    n = 5000
    ca, cb = c["pos"][:n].mean(0), c["pos"][n:].mean(0)
    ra = np.linalg.norm(c["pos"][:n] - ca, axis=1).max()
    rb = np.linalg.norm(c["pos"][n:] - cb, axis=1).max()
    assert abs(rb / ra - 0.5) < 0.03
    assert abs(c["mass"][n:].sum() / c["mass"][:n].sum() - 8.0) < 1e-3       # 8x the mass in 1/8 of the volume: 64x density
    assert c["vel"][:n, 0].mean() > 0 > c["vel"][n:, 0].mean()
    assert abs((cb - ca)[0] - 3.0 * ic.scaled_radius(n)) < 0.05 * ic.scaled_radius(n)


def test_rotating_sphere_is_rigid_rotation_about_z():
    from sphb200 import ic
    c = ic.make_rotating_sphere(4000, omega=0.02, seed=5)
    lz = (c["mass"] * (c["pos"][:, 0] * c["vel"][:, 1] - c["pos"][:, 1] * c["vel"][:, 0])).sum()
    inertia = (c["mass"] * (c["pos"][:, 0] ** 2 + c["pos"][:, 1] ** 2)).sum()
    assert abs(lz / inertia - 0.02) < 1e-5
    assert np.all(c["vel"][:, 2] == 0)


def test_polytrope_follows_the_hydrostatic_profile_of_the_reference_eos():
    from sphb200 import ic
    K, G, M, n = 1000.0, 1.0, 100.0, 200_000
    R = ic.polytrope_radius(K, G)
    assert abs(R - 39.633) < 1e-2                                      # pi sqrt(K / 2 pi G), SURVEY.md 8c
    c = ic.make_polytrope(n, total_mass=M, K=K, G=G, seed=9)
    r = np.linalg.norm(c["pos"].astype(np.float64), axis=1)
    assert r.max() <= R * (1 + 1e-6)
    # enclosed mass against the analytic m(r)/M = (sin xi - xi cos xi)/pi
    for frac in (0.2, 0.4, 0.6, 0.8, 0.95):
        xi = np.pi * frac
        want = (np.sin(xi) - xi * np.cos(xi)) / np.pi
        got = (r <= frac * R).mean()
        assert abs(got - want) < 4.0 * np.sqrt(want * (1 - want) / n) + 1e-4
    # isotropy and equal masses
    assert np.abs(c["pos"].mean(0)).max() < 0.2
    assert np.all(c["mass"] == np.float32(M / n))
    # h tracks the local density: ~50 particles inside 2h in the bulk
    rho_c = M * np.pi / (4 * R ** 3)
    inner = r < 0.3 * R
    xi = np.pi * r[inner] / R
    rho = rho_c * np.sin(xi) / xi
    nb = rho / (M / n) * 4.0 / 3.0 * np.pi * (2.0 * c["h"][inner].astype(np.float64)) ** 3
    assert abs(np.median(nb) - 50.0) < 0.5
    # hydrostatic balance of the analytic profile: dP/dr = -G m(r) rho / r^2 with P = K rho^2
    rr = np.linspace(0.05, 0.95, 50) * R
    x = np.pi * rr / R
    rho_a = rho_c * np.sin(x) / x
    drho = rho_c * (x * np.cos(x) - np.sin(x)) / x ** 2 * (np.pi / R)
    lhs = 2.0 * K * rho_a * drho
    m_r = M * (np.sin(x) - x * np.cos(x)) / np.pi
    rhs = -G * m_r * rho_a / rr ** 2
    np.testing.assert_allclose(lhs, rhs, rtol=1e-9)


def test_polytrope_group_velocity_and_spin():
    from sphb200 import ic
    c = ic.make_polytrope(20000, seed=2, center=(10.0, 0.0, -5.0), velocity=(0.5, 0.0, 0.0), omega=0.01)
    assert np.abs(c["pos"].mean(0) - np.array([10.0, 0.0, -5.0])).max() < 0.5
    assert abs(c["vel"][:, 0].mean() - 0.5) < 0.02
    rel = c["pos"].astype(np.float64) - np.array([10.0, 0.0, -5.0])
    vrel = c["vel"].astype(np.float64) - np.array([0.5, 0.0, 0.0])
    lz = (rel[:, 0] * vrel[:, 1] - rel[:, 1] * vrel[:, 0]).sum()
    inertia = (rel[:, 0] ** 2 + rel[:, 1] ** 2).sum()
    assert abs(lz / inertia - 0.01) < 1e-4
