"""CPU property tests of the two "threshold" arguments the CUDA path relies on (numpy float32 arithmetic is IEEE
round-to-nearest for *, /, sqrt, exactly what __fmul_rn / __fdiv_rn / __fsqrt_rn compute on the device).

1. Neighbor keep rule (csrc/ctx.cuh: sph_keep_threshold, used by k_cell_neighbors): the reference keeps the pair iff
       d2 < ((s*s)*2)*2 with s = max(h_i, h_j)      (SplineKernel.Interacts, A/Util/SplineKernel.cs:47-53)
   and Kernel(r, h_i) > 0 or Kernel(r, h_j) > 0, i.e. fsqrt_rn(d2) < 2 h_i or fsqrt_rn(d2) < 2 h_j   (:62, h < 1e5).
   Claim: this is d2 < max(C(h_i), C(h_j)) with C(h) = min(4 RN(h*h), t(h)), t(h) = smallest float whose rounded root
   reaches 2h.
2. Barnes-Hut MAC (csrc/kernels_tree.cu: mac_threshold): AcceptApproximation is RN(b_sq / r_sq) < theta^2
   (A/Systems/GravityFieldSystem.cs:229-247).  Claim: this is r_sq > T(b_sq) with T = largest float whose quotient is
   still >= theta^2.
"""
import numpy as np

f32 = np.float32


def _bits(x):
    return np.asarray(x, f32).view(np.uint32)


def _from_bits(u):
    return np.asarray(u, np.uint32).view(f32)


def keep_threshold(h):
    """numpy mirror of sph_keep_threshold (vectorised ulp stepping)."""
    h = np.asarray(h, f32)
    c = h * f32(2.0)
    u = _bits(c * c).astype(np.int64)
    for _ in range(8):                      # step down while the predecessor's root still reaches c
        down = (u > 0) & (np.sqrt(_from_bits(np.maximum(u - 1, 0))) >= c)
        if not down.any():
            break
        u = np.where(down, u - 1, u)
    for _ in range(8):                      # step up while the root falls short
        up = np.sqrt(_from_bits(u)) < c
        if not up.any():
            break
        u = np.where(up, u + 1, u)
    a = ((h * h) * f32(2.0)) * f32(2.0)
    return np.minimum(a, _from_bits(u))


def keep_literal(d2, hi, hj):
    s = np.maximum(hi, hj)
    first = d2 < ((s * s) * f32(2.0)) * f32(2.0)
    r = np.sqrt(d2)
    return first & ((r < hi * f32(2.0)) | (r < hj * f32(2.0)))


def test_keep_threshold_equals_literal_predicate_around_the_edge():
    rng = np.random.default_rng(0)
    n = 200_000
    hi = np.exp(rng.uniform(np.log(1e-3), np.log(9e4), n)).astype(f32)
    hj = (hi * np.exp(rng.uniform(-1.0, 1.0, n))).astype(f32)
    hj = np.minimum(hj, f32(9e4))
    ci, cj = keep_threshold(hi), keep_threshold(hj)
    cm = np.maximum(ci, cj)
    for off in range(-4, 5):                # d2 within 4 ulps of the threshold, both sides
        d2 = _from_bits((_bits(cm).astype(np.int64) + off).astype(np.uint32))
        np.testing.assert_array_equal(d2 < cm, keep_literal(d2, hi, hj), err_msg="offset %d ulps" % off)
    # and far from it
    d2 = (cm * rng.uniform(0.0, 2.0, n).astype(f32)).astype(f32)
    np.testing.assert_array_equal(d2 < cm, keep_literal(d2, hi, hj))


def test_keep_threshold_is_monotone_in_h():
    h = np.sort(np.exp(np.random.default_rng(1).uniform(np.log(1e-3), np.log(9e4), 100_000)).astype(f32))
    c = keep_threshold(h)
    assert np.all(np.diff(c) >= 0)          # max(C_i, C_j) == C(max(h_i, h_j)) needs this
    hn = _from_bits(_bits(h) + 1)           # neighbouring floats too
    assert np.all(keep_threshold(hn) >= c)


def mac_threshold(b_sq, theta2):
    """numpy mirror of mac_threshold."""
    b_sq = np.asarray(b_sq, f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = _bits(b_sq / theta2).astype(np.int64)
        for _ in range(8):
            down = (u > 0) & ~((b_sq / _from_bits(u)) >= theta2)
            if not down.any():
                break
            u = np.where(down, u - 1, u)
        for _ in range(8):
            up = (b_sq / _from_bits(u + 1)) >= theta2
            if not up.any():
                break
            u = np.where(up, u + 1, u)
    return np.where(b_sq > 0, _from_bits(u), f32(0.0))


def test_mac_threshold_equals_the_division_predicate():
    rng = np.random.default_rng(2)
    n = 200_000
    theta2 = f32(0.7) * f32(0.7)
    b_sq = np.exp(rng.uniform(np.log(1e-6), np.log(1e8), n)).astype(f32)
    b_sq[:100] = 0.0                        # single-particle nodes with point boxes
    T = mac_threshold(b_sq, theta2)
    with np.errstate(divide="ignore", invalid="ignore"):
        for off in range(-4, 5):
            r_sq = _from_bits(np.maximum(_bits(T).astype(np.int64) + off, 0).astype(np.uint32))
            literal = (b_sq / r_sq) < theta2    # 0/0 = NaN -> False: a target on the node centre rejects it
            np.testing.assert_array_equal(r_sq > T, literal, err_msg="offset %d ulps" % off)
        r_sq = (T * rng.uniform(0.0, 3.0, n).astype(f32)).astype(f32)
        np.testing.assert_array_equal(r_sq > T, (b_sq / r_sq) < theta2)


def test_sign_of_difference_decides_the_comparison():
    """k_tree_walk takes the accept bit from the sign of T - r_sq: distinct floats never subtract to zero."""
    rng = np.random.default_rng(3)
    T = np.exp(rng.uniform(np.log(1e-6), np.log(1e9), 200_000)).astype(f32)
    for off in range(-3, 4):
        r_sq = _from_bits((_bits(T).astype(np.int64) + off).astype(np.uint32))
        np.testing.assert_array_equal(np.signbit(T - r_sq), r_sq > T)
