// sph_oracle.cpp -- CPU ORACLE for the PlanetModel-SPH per-timestep hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing in the product path (the sphb200 CUDA library, its C ABI,
// the Python/C++ host mirrors) may link, import or call this file.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
//
// What it is: an op-for-op fp32 restatement of the reference's C#/Burst job path (the reference cannot be
// compiled here: no C#/Unity toolchain).  Each function cites the reference file:line it
// follows.  Paths: A/ = /root/reference/Assets/Scripts/, UP/ = /root/reference/UpstreamPackages/
// com.unity.physics@0.6.0-preview.3/Unity.Physics/.
//
// PARITY PINNING STATUS: the reference has *no* tests, fixtures or golden vectors for the SPH path
// (A/Util/SplineKernel.cs:43 "TODO: learn to write tests in unity!").  The SPH arithmetic is therefore
// "parity unpinned" against reference outputs; it is pinned instead by (i) the invariants the reference
// states in its own comments (SplineKernel.cs:29-43), (ii) closed forms, (iii) an independent numpy
// float32 restatement (tests/test_oracle_known_answers.py), and -- for the vendored Unity.Physics pieces that
// DO have reference tests -- (iv) the known answers in UT/PlayModeTests (MotionTests.cs:50-91 ExpandAabb /
// CalculateExpansion, SphereColliderTests.cs:81-118).
//
// Third-party arithmetic not in the tree (com.unity.mathematics 1.2.1) is restated as:
//   math.dot(a,a) = a.x*a.x + a.y*a.y + a.z*a.z (left to right), math.length = sqrtf(dot), math.max = fmaxf,
//   math.pow = correctly rounded powf (computed in double), Mathf.PI = 3.14159274f.
// No FMA contraction (compile with -ffp-contract=off), IEEE-RN '/' and sqrt.
//
// Build: g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).

#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

const float kPI = 3.14159274f;  // UnityEngine.Mathf.PI (fp32)

struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };

inline F3 sub3(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 ld3(const float* p, int64_t i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

// A/Util/SplineKernel.cs:44
inline float Kappa() { return 2.0f; }

// A/Util/SplineKernel.cs:47-53
inline bool Interacts(F3 r_i, F3 r_j, float size_i, float size_j) {
    F3 d = sub3(r_i, r_j);
    float size = fmaxf(size_i, size_j);
    float distanceSq = dot3(d, d);
    return distanceSq < size * size * Kappa() * Kappa();
}

// A/Util/SplineKernel.cs:55-89
inline float Kernel(float distance, float size) {
    float retval = 0.0f;
    if (distance >= size * Kappa()) return retval;
    float r_over_h = distance / size;
    float pi_h_cube = kPI * size * size * size;
    if (distance < size) {
        float r_over_h_sq = r_over_h * r_over_h;
        float numerator = 1.0f - 1.5f * r_over_h_sq + 0.75f * (r_over_h_sq * r_over_h);
        retval = numerator / pi_h_cube;
    } else {
        float inner_term = 2.0f - r_over_h;
        float numerator = inner_term * inner_term * inner_term;
        float denominator = 4.0f * pi_h_cube;
        retval = numerator / denominator;
    }
    return retval;
}

// A/Util/SplineKernel.cs:115-148.  fix_q1 = 0 reproduces the reference (quirk Q1: +3q in the inner branch).
inline float KernelDeriv(float distance, float size, int fix_q1) {
    float retval = 0.0f;
    if (distance >= size * Kappa()) return retval;
    float r_over_h = distance / size;
    float pi_h_4th = kPI * size * size * size * size;
    if (distance < size) {
        float r_over_h_sq = r_over_h * r_over_h;
        float lead = fix_q1 ? -3.0f : 3.0f;
        float numerator = lead * r_over_h + 2.25f * r_over_h_sq;
        retval = numerator / pi_h_4th;
    } else {
        float inner_term = 2.0f - r_over_h;
        float numerator = -3.0f * inner_term * inner_term;
        float denominator = 4.0f * pi_h_4th;
        retval = numerator / denominator;
    }
    return retval;
}

// A/Util/SplineKernel.cs:102-111
inline F4 KernelAndGradienti(F3 r_i, F3 r_j, float size, int fix_q1) {
    F3 d = sub3(r_i, r_j);
    float distance = sqrtf(dot3(d, d));
    float s = KernelDeriv(distance, size, fix_q1) / distance;  // r==0 -> 0/0 = NaN (quirk Q9)
    float kernel = Kernel(distance, size);
    return {d.x * s, d.y * s, d.z * s, kernel};
}

struct Interaction { F4 kthis; F4 ksym; };

// A/Systems/KernelSystem.cs:305-334
inline Interaction CalculateInteraction(F3 r_i, F3 r_j, float h_i, float h_j, int fix_q1) {
    F4 ki = KernelAndGradienti(r_i, r_j, h_i, fix_q1);
    F4 kj = KernelAndGradienti(r_i, r_j, h_j, fix_q1);
    Interaction it;
    it.kthis = ki;
    it.ksym = {(ki.x + kj.x) * 0.5f, (ki.y + kj.y) * 0.5f, (ki.z + kj.z) * 0.5f, (ki.w + kj.w) * 0.5f};
    return it;
}

// Effective neighbor rule = FilterPairs predicate (KernelSystem.cs:617) AND keep rule (KernelSystem.cs:269,283).
inline bool IsNeighbor(F3 r_i, F3 r_j, float h_i, float h_j) {
    if (!Interacts(r_i, r_j, h_i, h_j)) return false;
    Interaction it = CalculateInteraction(r_i, r_j, h_i, h_j, 0);
    return it.ksym.w > 0.0f;
}

// A/Systems/GravityFieldSystem.cs:332-356
inline F4 GravityContributionParticle(F3 r_i, F3 r_j, float m, float a, float G) {
    F3 d = sub3(r_i, r_j);
    float r = sqrtf(dot3(d, d));
    float grav_mag_over_r, grav_potential;
    if (r < a) {
        float x = r / a;
        float x_sq = x * x;
        float x_cube = x_sq * x;
        float x_5th = x_sq * x_cube;
        grav_mag_over_r = (m / (a * a * a)) * (8.0f - 9.0f * x + 2.0f * x_cube);
        grav_potential = -(m / a) * (2.4f - 4.0f * x_sq + 3.0f * x_cube - 0.4f * x_5th);
    } else {
        grav_mag_over_r = m / (r * r * r);
        grav_potential = -(m / r);
    }
    return {G * (d.x * grav_mag_over_r), G * (d.y * grav_mag_over_r), G * (d.z * grav_mag_over_r),
            G * grav_potential};
}

// A/Systems/GravityFieldSystem.cs:367-443  (GravitationalMoment)
struct Moment {
    F3 cm{0, 0, 0};
    float m = 0;
    // :398-411 Accumulate (P2M and M2M share it, :415-418)
    void Accumulate(F3 c, float first) {
        if (first != 0.0f) {
            float newMoment = m + first;
            F3 n;
            n.x = (cm.x * m + c.x * first) / newMoment;
            n.y = (cm.y * m + c.y * first) / newMoment;
            n.z = (cm.z * m + c.z * first) / newMoment;
            cm = n;
            m = newMoment;
        }
    }
    // :428-442 M2P
    F4 GravityContribution(F3 r_i, float G) const {
        F3 d = sub3(r_i, cm);
        float r = sqrtf(dot3(d, d));
        float g = m / (r * r * r);
        float phi = -(m / r);
        return {G * (d.x * g), G * (d.y * g), G * (d.z * g), G * phi};
    }
};

struct Aabb { F3 lo, hi; };

// A/Systems/GravityFieldSystem.cs:229-247
inline bool AcceptApproximation(F3 r_i, const Moment& mo, const Aabb& bv, float theta) {
    F3 d = sub3(r_i, mo.cm);
    float r_sq = dot3(d, d);
    F3 b;
    b.x = fmaxf(bv.hi.x - mo.cm.x, mo.cm.x - bv.lo.x);
    b.y = fmaxf(bv.hi.y - mo.cm.y, mo.cm.y - bv.lo.y);
    b.z = fmaxf(bv.hi.z - mo.cm.z, mo.cm.z - bv.lo.z);
    float bmax_sq = dot3(b, b);
    return bmax_sq / r_sq < theta * theta;
}

// UP/Dynamics/Motion/Motion.cs:142-146 (MotionExpansion.ExpandAabb)
inline Aabb ExpandAabb(Aabb a, F3 lin, float uni) {
    Aabb o;
    o.hi.x = fmaxf(a.hi.x, a.hi.x + lin.x) + uni;
    o.hi.y = fmaxf(a.hi.y, a.hi.y + lin.y) + uni;
    o.hi.z = fmaxf(a.hi.z, a.hi.z + lin.z) + uni;
    o.lo.x = fminf(a.lo.x, a.lo.x + lin.x) - uni;
    o.lo.y = fminf(a.lo.y, a.lo.y + lin.y) - uni;
    o.lo.z = fminf(a.lo.z, a.lo.z + lin.z) - uni;
    return o;
}

// Particle collider box as the broadphase prepares it (quirk Q2): sphere of radius 2h centred on x
// (Physics_SphereCollider.cs:129-137), swept by v*dt with zero angular part (Motion.cs:117-122),
// then Expand(margin = CollisionTolerance*0.5 = 0.05) (UP/Collision/World/Broadphase.cs:200, 757-761;
// CollisionWorld.cs:32).  mode 1 = plain point bounds (non-reference option).
inline Aabb ParticleBox(F3 x, float h, F3 v, float dt, int aabb_mode) {
    Aabb a;
    if (aabb_mode == 1) { a.lo = x; a.hi = x; return a; }
    float rad = h * Kappa();
    a.lo = {x.x - rad, x.y - rad, x.z - rad};
    a.hi = {x.x + rad, x.y + rad, x.z + rad};
    F3 lin = {v.x * dt, v.y * dt, v.z * dt};
    a = ExpandAabb(a, lin, 0.0f);
    const float margin = 0.1f * 0.5f;
    a.lo = {a.lo.x - margin, a.lo.y - margin, a.lo.z - margin};
    a.hi = {a.hi.x + margin, a.hi.y + margin, a.hi.z + margin};
    return a;
}

inline Aabb Union(const Aabb& a, const Aabb& b) {
    return {{fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z)},
            {fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z)}};
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Scalar entry points (known-answer tests)
// ------------------------------------------------------------------------------------------------
ORC_API float orc_kernel(float r, float h) { return Kernel(r, h); }
ORC_API float orc_kernel_deriv(float r, float h, int fix_q1) { return KernelDeriv(r, h, fix_q1); }
ORC_API int orc_interacts(const float* ri, const float* rj, float hi, float hj) {
    return Interacts(ld3(ri, 0), ld3(rj, 0), hi, hj) ? 1 : 0;
}
ORC_API int orc_is_neighbor(const float* ri, const float* rj, float hi, float hj) {
    return IsNeighbor(ld3(ri, 0), ld3(rj, 0), hi, hj) ? 1 : 0;
}
ORC_API void orc_kernel_and_gradient(const float* ri, const float* rj, float h, int fix_q1, float* out4) {
    F4 k = KernelAndGradienti(ld3(ri, 0), ld3(rj, 0), h, fix_q1);
    out4[0] = k.x; out4[1] = k.y; out4[2] = k.z; out4[3] = k.w;
}
// out8 = KernelThis(4), KernelSymmetric(4)
ORC_API void orc_interaction(const float* ri, const float* rj, float hi, float hj, int fix_q1, float* out8) {
    Interaction it = CalculateInteraction(ld3(ri, 0), ld3(rj, 0), hi, hj, fix_q1);
    memcpy(out8, &it.kthis, 16);
    memcpy(out8 + 4, &it.ksym, 16);
}
ORC_API void orc_gravity_pair(const float* ri, const float* rj, float m, float a, float G, float* out4) {
    F4 g = GravityContributionParticle(ld3(ri, 0), ld3(rj, 0), m, a, G);
    memcpy(out4, &g, 16);
}
// moment accumulate: cm_m[4] in/out
ORC_API void orc_moment_accumulate(float* cm_m, const float* c, float m) {
    Moment mo; mo.cm = {cm_m[0], cm_m[1], cm_m[2]}; mo.m = cm_m[3];
    mo.Accumulate(ld3(c, 0), m);
    cm_m[0] = mo.cm.x; cm_m[1] = mo.cm.y; cm_m[2] = mo.cm.z; cm_m[3] = mo.m;
}
ORC_API void orc_moment_m2p(const float* cm_m, const float* ri, float G, float* out4) {
    Moment mo; mo.cm = {cm_m[0], cm_m[1], cm_m[2]}; mo.m = cm_m[3];
    F4 g = mo.GravityContribution(ld3(ri, 0), G);
    memcpy(out4, &g, 16);
}
ORC_API int orc_accept(const float* ri, const float* cm_m, const float* lo, const float* hi, float theta) {
    Moment mo; mo.cm = {cm_m[0], cm_m[1], cm_m[2]}; mo.m = cm_m[3];
    Aabb b{ld3(lo, 0), ld3(hi, 0)};
    return AcceptApproximation(ld3(ri, 0), mo, b, theta) ? 1 : 0;
}
// UP/Dynamics/Motion/Motion.cs:117-122, 142-146 -- pinned by UT MotionTests.cs:50-91
ORC_API void orc_expand_aabb(const float* lo, const float* hi, const float* lin, float uni, float* out_lo, float* out_hi) {
    Aabb o = ExpandAabb({ld3(lo, 0), ld3(hi, 0)}, ld3(lin, 0), uni);
    memcpy(out_lo, &o.lo, 12); memcpy(out_hi, &o.hi, 12);
}
ORC_API void orc_calculate_expansion(const float* linvel, const float* angvel, float angfactor, float dt, float* out_lin3, float* out_uni) {
    out_lin3[0] = linvel[0] * dt; out_lin3[1] = linvel[1] * dt; out_lin3[2] = linvel[2] * dt;
    F3 w = ld3(angvel, 0);
    *out_uni = fminf(sqrtf(dot3(w, w)) * dt * angfactor, angfactor);
}
ORC_API void orc_particle_box(const float* x, float h, const float* v, float dt, int aabb_mode, float* out_lo, float* out_hi) {
    Aabb o = ParticleBox(ld3(x, 0), h, ld3(v, 0), dt, aabb_mode);
    memcpy(out_lo, &o.lo, 12); memcpy(out_hi, &o.hi, 12);
}

// ------------------------------------------------------------------------------------------------
// Smoothing-length controller   A/Systems/ParticleSmoothingSystem.cs:21-86
// ------------------------------------------------------------------------------------------------
// radius_ratio = math.pow(TARGET/(float)n, 1.0f/3.0f).  Unity.Mathematics/Burst pow is not in the tree;
// restated as the correctly rounded fp32 power (double pow, one rounding).
ORC_API float orc_radius_ratio(float target, int n) {
    float ratio = target / (float)n;
    return (float)pow((double)ratio, (double)(1.0f / 3.0f));
}
ORC_API void orc_smoothing_update(int64_t n, const float* h_prev, const int* n_own, float target, float* h_next) {
    for (int64_t i = 0; i < n; i++) {
        float h = h_prev[i];
        if (n_own[i] != 0) {
            float rr = orc_radius_ratio(target, n_own[i]);
            h = h_prev[i] * 0.5f * (1.0f + rr);
        }
        h_next[i] = h;
    }
}

// ------------------------------------------------------------------------------------------------
// Neighbor sets (ground truth: brute force; cell list for sizes brute force cannot reach).
// Lists are CSR, per-particle neighbors in ascending index order (the canonical order; the reference's
// BVH emission order -- KernelSystem.cs:262-287 -- is not reproducible by any other structure).
// Returns total entries; if > cap nothing past cap is written (caller re-calls with a larger buffer).
// ------------------------------------------------------------------------------------------------
ORC_API int64_t orc_neighbors_brute(int64_t n, const float* pos, const float* h, int64_t* offsets, int32_t* nbr, int64_t cap) {
    std::vector<int32_t> cnt(n, 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        F3 ri = ld3(pos, i); int c = 0;
        for (int64_t j = 0; j < n; j++)
            if (j != i && IsNeighbor(ri, ld3(pos, j), h[i], h[j])) c++;
        cnt[i] = c;
    }
    offsets[0] = 0;
    for (int64_t i = 0; i < n; i++) offsets[i + 1] = offsets[i] + cnt[i];
    if (offsets[n] > cap) return offsets[n];
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        F3 ri = ld3(pos, i); int64_t o = offsets[i];
        for (int64_t j = 0; j < n; j++)
            if (j != i && IsNeighbor(ri, ld3(pos, j), h[i], h[j])) nbr[o++] = (int32_t)j;
    }
    return offsets[n];
}

// Cell list with cell edge >= 2*h_max*(1+1e-3): every pair that passes Interacts lies in adjacent cells.
// (Independent of the GPU's grid: this is the ground truth for larger N, itself checked against brute force.)
ORC_API int64_t orc_neighbors_grid(int64_t n, const float* pos, const float* h, int64_t* offsets, int32_t* nbr, int64_t cap) {
    if (n == 0) { offsets[0] = 0; return 0; }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}; double hmax = 0;
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], (double)pos[3 * i + k]); hi[k] = std::max(hi[k], (double)pos[3 * i + k]); }
        hmax = std::max(hmax, (double)h[i]);
    }
    double cell = 2.0 * hmax * 1.001; if (!(cell > 0)) cell = 1.0;
    int64_t dim[3];
    for (int k = 0; k < 3; k++) { dim[k] = (int64_t)std::floor((hi[k] - lo[k]) / cell) + 1; }
    while ((double)dim[0] * dim[1] * dim[2] > 64e6) { cell *= 1.26; for (int k = 0; k < 3; k++) dim[k] = (int64_t)std::floor((hi[k] - lo[k]) / cell) + 1; }
    int64_t ncell = dim[0] * dim[1] * dim[2];
    std::vector<int64_t> cid(n);
    std::vector<int64_t> cstart(ncell + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        int64_t c[3];
        for (int k = 0; k < 3; k++) { c[k] = (int64_t)std::floor(((double)pos[3 * i + k] - lo[k]) / cell); c[k] = std::min(std::max(c[k], (int64_t)0), dim[k] - 1); }
        cid[i] = (c[2] * dim[1] + c[1]) * dim[0] + c[0];
        cstart[cid[i] + 1]++;
    }
    for (int64_t c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
    std::vector<int32_t> members(n);
    { std::vector<int64_t> fill(cstart.begin(), cstart.end() - 1);
      for (int64_t i = 0; i < n; i++) members[fill[cid[i]]++] = (int32_t)i; }  // ascending i inside each cell
    auto gather = [&](int64_t i, std::vector<int32_t>& out) {
        out.clear();
        int64_t c = cid[i]; int64_t cx = c % dim[0], cy = (c / dim[0]) % dim[1], cz = c / (dim[0] * dim[1]);
        F3 ri = ld3(pos, i);
        for (int64_t z = std::max<int64_t>(cz - 1, 0); z <= std::min(cz + 1, dim[2] - 1); z++)
            for (int64_t y = std::max<int64_t>(cy - 1, 0); y <= std::min(cy + 1, dim[1] - 1); y++)
                for (int64_t x = std::max<int64_t>(cx - 1, 0); x <= std::min(cx + 1, dim[0] - 1); x++) {
                    int64_t cc = (z * dim[1] + y) * dim[0] + x;
                    for (int64_t k = cstart[cc]; k < cstart[cc + 1]; k++) {
                        int32_t j = members[k];
                        if (j != i && IsNeighbor(ri, ld3(pos, j), h[i], h[j])) out.push_back(j);
                    }
                }
        std::sort(out.begin(), out.end());
    };
    std::vector<int32_t> cnt(n);
#pragma omp parallel
    { std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 256)
      for (int64_t i = 0; i < n; i++) { gather(i, tmp); cnt[i] = (int32_t)tmp.size(); } }
    offsets[0] = 0;
    for (int64_t i = 0; i < n; i++) offsets[i + 1] = offsets[i] + cnt[i];
    if (offsets[n] > cap) return offsets[n];
#pragma omp parallel
    { std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 256)
      for (int64_t i = 0; i < n; i++) { gather(i, tmp); std::copy(tmp.begin(), tmp.end(), nbr + offsets[i]); } }
    return offsets[n];
}

// Materialised interaction buffer (DynamicBuffer<ParticleInteraction>, A/Components/Kernel.cs:5-16) for the
// entries of one CSR list: kthis[4*e], ksym[4*e].
ORC_API void orc_interactions(int64_t n, const float* pos, const float* h, const int64_t* offsets, const int32_t* nbr,
                              int fix_q1, float* kthis, float* ksym) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++)
        for (int64_t e = offsets[i]; e < offsets[i + 1]; e++) {
            int32_t j = nbr[e];
            Interaction it = CalculateInteraction(ld3(pos, i), ld3(pos, j), h[i], h[j], fix_q1);
            memcpy(kthis + 4 * e, &it.kthis, 16);
            memcpy(ksym + 4 * e, &it.ksym, 16);
        }
}

// ------------------------------------------------------------------------------------------------
// Density  A/Systems/DensityFieldSystem.cs:38-56 ; own-support count ParticleSmoothingSystem.cs:33-43
// ------------------------------------------------------------------------------------------------
ORC_API void orc_density(int64_t n, const float* pos, const float* h, const float* m, const int64_t* offsets,
                         const int32_t* nbr, float* rho, int32_t* n_own) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        float density = m[i] * Kernel(0.0f, h[i]);
        int own = 0;
        for (int64_t e = offsets[i]; e < offsets[i + 1]; e++) {
            int32_t j = nbr[e];
            Interaction it = CalculateInteraction(ld3(pos, i), ld3(pos, j), h[i], h[j], 0);
            density += m[j] * it.ksym.w;
            if (it.kthis.w > 0.0f) own++;
        }
        rho[i] = density;
        if (n_own) n_own[i] = own;
    }
}

// EOS  A/Systems/PressureFieldSystem.cs:30-34
ORC_API void orc_eos(int64_t n, const float* rho, float K, float* P) {
    for (int64_t i = 0; i < n; i++) P[i] = K * rho[i] * rho[i];
}

// Pressure gradient  A/Systems/PressureFieldSystem.cs:44-70
ORC_API void orc_pressure_grad(int64_t n, const float* pos, const float* h, const float* m, const float* rho,
                               const float* P, const int64_t* offsets, const int32_t* nbr, int fix_q1, float* gradP) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        F3 g{0, 0, 0};
        for (int64_t e = offsets[i]; e < offsets[i + 1]; e++) {
            int32_t j = nbr[e];
            Interaction it = CalculateInteraction(ld3(pos, i), ld3(pos, j), h[i], h[j], fix_q1);
            float s = m[j] / rho[j] * P[j];
            g.x += it.ksym.x * s; g.y += it.ksym.y * s; g.z += it.ksym.z * s;
        }
        gradP[3 * i] = g.x; gradP[3 * i + 1] = g.y; gradP[3 * i + 2] = g.z;
    }
}

// ------------------------------------------------------------------------------------------------
// Direct gravity  A/Systems/GravityFieldSystem.cs:249-303 (targets i0..i1, all sources, self skipped).
// accum_double = 0: literal fp32 sequential sum in index order (what the reference computes);
//              = 1: same per-pair fp32 arithmetic, accumulated in double (summation-order-free yardstick).
// ------------------------------------------------------------------------------------------------
ORC_API void orc_gravity_direct(int64_t n, const float* pos, const float* h, const float* m, float G,
                                int64_t i0, int64_t i1, int accum_double, float* grav4) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = i0; i < i1; i++) {
        F3 ri = ld3(pos, i); float a = h[i];
        if (accum_double) {
            double s[4] = {0, 0, 0, 0};
            for (int64_t j = 0; j < n; j++) {
                if (j == i) continue;
                F4 c = GravityContributionParticle(ri, ld3(pos, j), m[j], a, G);
                s[0] += c.x; s[1] += c.y; s[2] += c.z; s[3] += c.w;
            }
            for (int k = 0; k < 4; k++) grav4[4 * (i - i0) + k] = (float)s[k];
        } else {
            F4 s{0, 0, 0, 0};
            for (int64_t j = 0; j < n; j++) {
                if (j == i) continue;
                F4 c = GravityContributionParticle(ri, ld3(pos, j), m[j], a, G);
                s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
            }
            memcpy(grav4 + 4 * (i - i0), &s, 16);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NON-REFERENCE OPTION (roadmap A README.md:75-77 "Gravity kernel which conserves energy, see Price & Monaghan 2007"):
// gravity softened with the cubic-spline kernel of support 2h, symmetrised in the two smoothing lengths,
//     grad Phi_i = G sum_j m_j (r_i - r_j)/r * 0.5 [phi'(r,h_i) + phi'(r,h_j)],  Phi_i = G sum_j m_j 0.5 [phi(r,h_i) + phi(r,h_j)]
// with phi, phi' of P&M 2007 appendix A (Newtonian beyond r = 2h).  Evaluated in double: this is the yardstick of
// SPH_FLAG_PM07_SOFTENING, which the GPU builds as "main gravity kernel + neighbor-list correction" in fp32.
// ------------------------------------------------------------------------------------------------
namespace {
inline void Pm07Kernel(double r, double h, double& dphi_over_r, double& phi) {
    double q = r / h;
    if (q < 1.0) {
        dphi_over_r = (4.0 / 3.0 - 1.2 * q * q + 0.5 * q * q * q) / (h * h * h);
        phi = (2.0 / 3.0 * q * q - 0.3 * q * q * q * q + 0.1 * q * q * q * q * q - 1.4) / h;
    } else if (q < 2.0) {
        dphi_over_r = (8.0 / 3.0 - 3.0 * q + 1.2 * q * q - q * q * q / 6.0 - 1.0 / (15.0 * q * q * q)) / (h * h * h);
        phi = (4.0 / 3.0 * q * q - q * q * q + 0.3 * q * q * q * q - q * q * q * q * q / 30.0 - 1.6 + 1.0 / (15.0 * q)) / h;
    } else {
        dphi_over_r = 1.0 / (r * r * r);
        phi = -1.0 / r;
    }
}
}  // namespace

ORC_API void orc_pm07_kernel(double r, double h, double* out2) { Pm07Kernel(r, h, out2[0], out2[1]); }

// all sources, targets i0..i1, double accumulation
ORC_API void orc_gravity_direct_pm07(int64_t n, const float* pos, const float* h, const float* m, float G,
                                     int64_t i0, int64_t i1, float* grav4) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = i0; i < i1; i++) {
        double s[4] = {0, 0, 0, 0};
        for (int64_t j = 0; j < n; j++) {
            if (j == i) continue;
            double d[3] = {(double)pos[3 * i] - pos[3 * j], (double)pos[3 * i + 1] - pos[3 * j + 1], (double)pos[3 * i + 2] - pos[3 * j + 2]};
            double r = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (!(r > 0.0)) { s[3] += m[j] * 0.5 * (-1.4 / h[i] - 1.4 / h[j]); continue; }   // coincident: no force, phi(0)
            double fi, pi, fj, pj;
            Pm07Kernel(r, h[i], fi, pi); Pm07Kernel(r, h[j], fj, pj);
            double f = 0.5 * (fi + fj) * m[j];
            s[0] += d[0] * f; s[1] += d[1] * f; s[2] += d[2] * f; s[3] += m[j] * 0.5 * (pi + pj);
        }
        for (int k = 0; k < 4; k++) grav4[4 * (i - i0) + k] = (float)((double)G * s[k]);
    }
}

// Correction over the neighbor lists that turns a reference-law sum into the PM07 sum (what the tree path adds):
// out4[i] = G sum_{j in list(i)} m_j { PM07 pair - GravityContributionParticle pair }, double accumulation.
ORC_API void orc_gravity_pm07_correction(int64_t n, const float* pos, const float* h, const float* m, float G,
                                         const int64_t* offsets, const int32_t* nbr, float* out4) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        double s[4] = {0, 0, 0, 0};
        for (int64_t e = offsets[i]; e < offsets[i + 1]; e++) {
            int64_t j = nbr[e];
            double d[3] = {(double)pos[3 * i] - pos[3 * j], (double)pos[3 * i + 1] - pos[3 * j + 1], (double)pos[3 * i + 2] - pos[3 * j + 2]};
            double r = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            F4 ref = GravityContributionParticle(ld3(pos, i), ld3(pos, j), m[j], h[i], 1.0f);
            double fi = 0, pi = -1.4 / h[i], fj = 0, pj = -1.4 / h[j];
            if (r > 0.0) { Pm07Kernel(r, h[i], fi, pi); Pm07Kernel(r, h[j], fj, pj); }
            double f = 0.5 * (fi + fj) * m[j];
            s[0] += d[0] * f - ref.x; s[1] += d[1] * f - ref.y; s[2] += d[2] * f - ref.z; s[3] += m[j] * 0.5 * (pi + pj) - ref.w;
        }
        for (int k = 0; k < 4; k++) out4[4 * i + k] = (float)((double)G * s[k]);
    }
}

// ------------------------------------------------------------------------------------------------
// Integration: x += v*dt  (UP/Dynamics/Integrator/Integrator.cs:98-101, old v), then
//              v += (-gradP/rho - gradPhi)*dt  (A/Systems/VelocitySystem.cs:24-36)
// ------------------------------------------------------------------------------------------------
ORC_API void orc_integrate2(int64_t n, float* pos, float* vel, const float* rho, const float* gradP,
                            const float* grav4, float dt, int kick_drift) {
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            float v = vel[3 * i + k];
            if (!kick_drift) pos[3 * i + k] = pos[3 * i + k] + v * dt;
            float dvdt = -gradP[3 * i + k] / rho[i] - grav4[4 * i + k];
            float vn = v + dvdt * dt;
            vel[3 * i + k] = vn;
            if (kick_drift) pos[3 * i + k] = pos[3 * i + k] + vn * dt;   // non-reference option (README.md:90-93 roadmap)
        }
}
ORC_API void orc_integrate(int64_t n, float* pos, float* vel, const float* rho, const float* gradP,
                           const float* grav4, float dt) {
    orc_integrate2(n, pos, vel, rho, gradP, grav4, dt, 0);
}

// ------------------------------------------------------------------------------------------------
// Spatial keys.  This part has no reference counterpart (it replaces the Unity.Physics broadphase); it is
// the *specification* the CUDA key/sort kernels must match bit-for-bit ("sort orders bit-exact").
// GridParams layout is shared with include/sphb200.h (sph_GridParams).
// ------------------------------------------------------------------------------------------------
struct GridParams {
    float min[3];      // global AABB min of positions
    float cell;        // cell edge
    float fine_scale;  // 1024 / (cell * 2^bits)
    int32_t bits;      // cells per axis = 2^bits
    float hmax;
    float ext;         // max extent over the three axes
    float href;        // typical h (mean of the fp32 bit patterns, reinterpreted)
    int32_t stencil;   // S
};

static inline uint32_t expand10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

static const int kStencilMax = 4;

// Cell edge: 2.002 x a *typical* h (not h_max, which would inflate every cell by the tail of the h distribution);
// pairs with a large h are covered by a stencil of S cells per axis, S*cell >= 2.002*h_max, S <= 4.
// The typical h must be order-independent to be reproducible on the GPU: it is the float whose bit pattern is the
// integer mean of all h bit patterns (monotone in h, between the geometric and arithmetic means).
ORC_API void orc_grid_params(int64_t n, const float* pos, const float* h, int max_bits, GridParams* g) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}; float hmax = 0.0f;
    uint64_t bitsum = 0;
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], pos[3 * i + k]); hi[k] = fmaxf(hi[k], pos[3 * i + k]); }
        hmax = fmaxf(hmax, h[i]);
        uint32_t u; memcpy(&u, &h[i], 4); bitsum += u;
    }
    uint32_t ub = n > 0 ? (uint32_t)(bitsum / (uint64_t)n) : 0u;
    float href; memcpy(&href, &ub, 4);
    float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    float reach = hmax * 2.002f;
    float cell = href * 2.002f;
    if (!(cell * (float)kStencilMax >= reach)) cell = reach / (float)kStencilMax;
    int bits = 0;
    for (; bits < max_bits; bits++)
        if (cell * (float)(1 << bits) > ext) break;
    if (!(cell * (float)(1 << bits) > ext)) cell = (ext * 1.0001f) / (float)(1 << bits);
    if (!(cell > 0.0f)) cell = 1.0f;
    int S = 1;
    while (S < kStencilMax && !(cell * (float)S >= reach)) S++;
    for (int k = 0; k < 3; k++) g->min[k] = lo[k];
    g->cell = cell; g->bits = bits; g->hmax = hmax; g->ext = ext; g->href = href; g->stencil = S;
    g->fine_scale = 1024.0f / (cell * (float)(1 << bits));
}

ORC_API void orc_morton_keys(int64_t n, const float* pos, const GridParams* g, uint32_t* keys) {
    for (int64_t i = 0; i < n; i++) {
        uint32_t q[3];
        for (int k = 0; k < 3; k++) {
            float s = (pos[3 * i + k] - g->min[k]) * g->fine_scale;
            int v = (int)s;  // truncation; s >= 0
            if (v < 0) v = 0; if (v > 1023) v = 1023;
            q[k] = (uint32_t)v;
        }
        keys[i] = expand10(q[0]) | (expand10(q[1]) << 1) | (expand10(q[2]) << 2);
    }
}

// stable ascending sort of (key, original index): order[k] = original index at sorted slot k
ORC_API void orc_sort_order(int64_t n, const uint32_t* keys, uint32_t* order) {
    std::iota(order, order + n, 0u);
    std::stable_sort(order, order + n, [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
}

// ------------------------------------------------------------------------------------------------
// LBVH over Morton-sorted particles (Karras 2012 binary radix tree) -- the structure that replaces the
// Unity.Physics 4-ary BVH (UP/Collision/Geometry/BoundingVolumeHierarchy.cs:40-83) as the Barnes-Hut tree.
// Node ids: internal 0..n-2 (root 0), leaf particle s -> n-1+s.  All arrays sized 2n-1.
// A node whose particle range holds <= leaf_max particles plays the role of a reference *leaf node*
// (<= 4 bodies, BoundingVolumeHierarchy.cs:57): it is never descended, its bodies are summed directly.
// ------------------------------------------------------------------------------------------------
static inline int delta_fn(const uint32_t* keys, int64_t n, int64_t i, int64_t j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __builtin_clz((uint32_t)i ^ (uint32_t)j);  // i != j here
    return __builtin_clz(a ^ b);
}

ORC_API void orc_lbvh_topology(int64_t n, const uint32_t* keys, int32_t* left, int32_t* right, int32_t* parent,
                               int32_t* first, int32_t* last) {
    // leaves
    for (int64_t s = 0; s < n; s++) { first[n - 1 + s] = (int32_t)s; last[n - 1 + s] = (int32_t)s; left[n - 1 + s] = -1; right[n - 1 + s] = -1; }
    if (n == 1) { parent[0] = -1; return; }
    parent[0] = -1;
    for (int64_t i = 0; i < n - 1; i++) {
        int d = (delta_fn(keys, n, i, i + 1) - delta_fn(keys, n, i, i - 1)) > 0 ? 1 : -1;
        int dmin = delta_fn(keys, n, i, i - d);
        int64_t lmax = 2;
        while (delta_fn(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
        int64_t l = 0;
        for (int64_t t = lmax / 2; t >= 1; t /= 2)
            if (delta_fn(keys, n, i, i + (l + t) * d) > dmin) l += t;
        int64_t j = i + l * d;
        int dnode = delta_fn(keys, n, i, j);
        int64_t s = 0;
        for (int64_t t = (l + 1) / 2;; t = (t + 1) / 2) {
            if (delta_fn(keys, n, i, i + (s + t) * d) > dnode) s += t;
            if (t == 1) break;
        }
        int64_t gamma = i + s * d + std::min(d, 0);
        int64_t lo = std::min(i, j), hi = std::max(i, j);
        int32_t lc = (lo == gamma) ? (int32_t)(n - 1 + gamma) : (int32_t)gamma;
        int32_t rc = (hi == gamma + 1) ? (int32_t)(n - 1 + gamma + 1) : (int32_t)(gamma + 1);
        left[i] = lc; right[i] = rc; parent[lc] = (int32_t)i; parent[rc] = (int32_t)i;
        first[i] = (int32_t)lo; last[i] = (int32_t)hi;
    }
}

// Moments + boxes for every node.  pos/vel/h/m are in *sorted* order.
//  - nodes with count <= leaf_max: reference leaf rule -- running mass-weighted mean over the bodies in
//    Data (= sorted) order from a zero moment (GravityFieldSystem.cs:398-411, 485-500)
//  - larger nodes: Accumulate(left child), Accumulate(right child) (GravityFieldSystem.cs:513-521)
//  - box = union of the particles' collider boxes (Q2) or points (aabb_mode 1); min/max is order-free.
// mom4[4*node] = (cm.xyz, M); lo3/hi3[3*node].
ORC_API void orc_lbvh_moments(int64_t n, const float* pos, const float* vel, const float* h, const float* m,
                              const int32_t* left, const int32_t* right, const int32_t* first, const int32_t* last,
                              int leaf_max, int aabb_mode, float dt, float* mom4, float* lo3, float* hi3) {
    int64_t nn = 2 * n - 1;
    std::vector<char> done(nn, 0);
    // small nodes directly
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t k = 0; k < nn; k++) {
        int cnt = last[k] - first[k] + 1;
        if (cnt > leaf_max) continue;
        Moment mo; Aabb b{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
        for (int s = first[k]; s <= last[k]; s++) {
            mo.Accumulate(ld3(pos, s), m[s]);
            b = Union(b, ParticleBox(ld3(pos, s), h[s], ld3(vel, s), dt, aabb_mode));
        }
        mom4[4 * k] = mo.cm.x; mom4[4 * k + 1] = mo.cm.y; mom4[4 * k + 2] = mo.cm.z; mom4[4 * k + 3] = mo.m;
        memcpy(lo3 + 3 * k, &b.lo, 12); memcpy(hi3 + 3 * k, &b.hi, 12);
        done[k] = 1;
    }
    // large nodes: post-order with an explicit stack (GravityFieldSystem.cs:465-539 pattern)
    if (n >= 1 && !done[0]) {
        std::vector<int32_t> stack; stack.push_back(0);
        while (!stack.empty()) {
            int32_t k = stack.back();
            int32_t l = left[k], r = right[k];
            if (done[l] && done[r]) {
                stack.pop_back();
                Moment mo;
                mo.Accumulate({mom4[4 * l], mom4[4 * l + 1], mom4[4 * l + 2]}, mom4[4 * l + 3]);
                mo.Accumulate({mom4[4 * r], mom4[4 * r + 1], mom4[4 * r + 2]}, mom4[4 * r + 3]);
                mom4[4 * k] = mo.cm.x; mom4[4 * k + 1] = mo.cm.y; mom4[4 * k + 2] = mo.cm.z; mom4[4 * k + 3] = mo.m;
                for (int c = 0; c < 3; c++) { lo3[3 * k + c] = fminf(lo3[3 * l + c], lo3[3 * r + c]); hi3[3 * k + c] = fmaxf(hi3[3 * l + c], hi3[3 * r + c]); }
                done[k] = 1;
            } else {
                if (!done[l]) stack.push_back(l);
                if (!done[r]) stack.push_back(r);
            }
        }
    }
}

// Tree walk  A/Systems/GravityFieldSystem.cs:133-215 on the LBVH above.
// Per target: explicit stack, pop; Accept -> M2P, numApprox++; else leaf (count<=leaf_max) -> P2P over its
// bodies INCLUDING the target itself (quirk Q3), numParticles++ each; else push children (left then right;
// LIFO pops right first, matching the reference's "push in Data order" :201-206).
// Targets t0..t1 (sorted indices).  accum_double as in orc_gravity_direct.
ORC_API void orc_tree_walk(int64_t n, const float* pos, const float* h, const float* m,
                           const int32_t* left, const int32_t* right, const int32_t* first, const int32_t* last,
                           const float* mom4, const float* lo3, const float* hi3,
                           int leaf_max, float theta, float G, int64_t t0, int64_t t1, int accum_double,
                           float* grav4, int32_t* num_particles, int32_t* num_approx) {
#pragma omp parallel
    {
        std::vector<int32_t> stack; stack.reserve(256);
#pragma omp for schedule(dynamic, 64)
        for (int64_t t = t0; t < t1; t++) {
            F3 ri = ld3(pos, t); float a = h[t];
            F4 g{0, 0, 0, 0}; double gd[4] = {0, 0, 0, 0};
            int np = 0, na = 0;
            stack.clear(); stack.push_back(0);
            do {
                int32_t k = stack.back(); stack.pop_back();
                Moment mo; mo.cm = {mom4[4 * k], mom4[4 * k + 1], mom4[4 * k + 2]}; mo.m = mom4[4 * k + 3];
                Aabb b{ld3(lo3, k), ld3(hi3, k)};
                if (AcceptApproximation(ri, mo, b, theta)) {
                    F4 c = mo.GravityContribution(ri, G);
                    if (accum_double) { gd[0] += c.x; gd[1] += c.y; gd[2] += c.z; gd[3] += c.w; }
                    else { g.x += c.x; g.y += c.y; g.z += c.z; g.w += c.w; }
                    na++;
                } else if (last[k] - first[k] + 1 <= leaf_max) {
                    for (int s = first[k]; s <= last[k]; s++) {
                        F4 c = GravityContributionParticle(ri, ld3(pos, s), m[s], a, G);
                        if (accum_double) { gd[0] += c.x; gd[1] += c.y; gd[2] += c.z; gd[3] += c.w; }
                        else { g.x += c.x; g.y += c.y; g.z += c.z; g.w += c.w; }
                        np++;
                    }
                } else {
                    stack.push_back(left[k]); stack.push_back(right[k]);
                }
            } while (!stack.empty());
            if (accum_double) { g.x = (float)gd[0]; g.y = (float)gd[1]; g.z = (float)gd[2]; g.w = (float)gd[3]; }
            memcpy(grav4 + 4 * (t - t0), &g, 16);
            if (num_particles) num_particles[t - t0] = np;
            if (num_approx) num_approx[t - t0] = na;
        }
    }
}

ORC_API int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_num_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

// ------------------------------------------------------------------------------------------------
// Reference-SHAPED tree: a 4-ary BVH built top-down over the centres of the particles'
// collider boxes the way Unity.Physics' builder does it -- large ranges: split at the spatial median of the longest
// axis, then each half again on its own longest axis (ProcessLargeRange, BoundingVolumeHierarchyBuilder.cs:331-345);
// ranges of <= 32 points: sort along the longest axis and peel leaves of 4 (ProcessSmallRange, :294-329); leaves hold
// <= 4 bodies; node boxes are unions of the children's collider boxes (Refit, :558-600).  It is NOT bit-faithful to the
// vendored builder (quirk Q10 and the multi-threaded branch numbering are not reproduced: tree shape is not a parity
// target); it exists to show that the LBVH walk and a Unity-shaped walk sit in the same accuracy envelope.
// The moment pass and the walk are the same reference arithmetic as above (GravityFieldSystem.cs:133-215, 465-539).
// ------------------------------------------------------------------------------------------------
namespace {
struct Node4 {
    int child[4];   // internal: node ids; leaf: body ids; -1 = unused
    int nchild = 0;
    bool leaf = false;
    Aabb box;
    Moment mom;
};
struct Tree4 {
    std::vector<Node4> nodes;
    const float* cen;   // box centres (3 per body)
    std::vector<int> idx;
    int longest_axis(int s, int len, float& mid) const {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int k = s; k < s + len; k++)
            for (int c = 0; c < 3; c++) { lo[c] = fminf(lo[c], cen[3 * idx[k] + c]); hi[c] = fmaxf(hi[c], cen[3 * idx[k] + c]); }
        int ax = 0;
        for (int c = 1; c < 3; c++) if (hi[c] - lo[c] > hi[ax] - lo[ax]) ax = c;
        mid = 0.5f * (lo[ax] + hi[ax]);
        return ax;
    }
    // spatial-median split of [s, s+len); falls back to halving by count when a side would hold < min_items
    int split(int s, int len, int min_items) {
        float mid; int ax = longest_axis(s, len, mid);
        auto it = std::partition(idx.begin() + s, idx.begin() + s + len, [&](int b) { return cen[3 * b + ax] < mid; });
        int l = (int)(it - (idx.begin() + s));
        if (l < min_items || len - l < min_items) {
            std::sort(idx.begin() + s, idx.begin() + s + len, [&](int a, int b) { return cen[3 * a + ax] < cen[3 * b + ax]; });
            l = len / 2;
        }
        return l;
    }
    int make_leaf(int s, int len) {
        Node4 nd; nd.leaf = true; nd.nchild = len;
        for (int k = 0; k < 4; k++) nd.child[k] = k < len ? idx[s + k] : -1;
        nodes.push_back(nd);
        return (int)nodes.size() - 1;
    }
    int build(int s, int len) {
        if (len <= 4) return make_leaf(s, len);
        int self = (int)nodes.size();
        nodes.emplace_back();
        int sub_s[4], sub_l[4], ns = 0;
        if (len <= 32) {
            float mid; int ax = longest_axis(s, len, mid);
            std::sort(idx.begin() + s, idx.begin() + s + len, [&](int a, int b) { return cen[3 * a + ax] < cen[3 * b + ax]; });
            int cs = s, cl = len;
            while (cl > 4 && ns < 3) { sub_s[ns] = cs; sub_l[ns] = 4; ns++; cs += 4; cl -= 4; }
            if (cl > 0) { sub_s[ns] = cs; sub_l[ns] = cl; ns++; }
        } else {
            int l = split(s, len, 2);
            int ll = split(s, l, 1), rl = split(s + l, len - l, 1);
            int ss[4] = {s, s + ll, s + l, s + l + rl}, sl[4] = {ll, l - ll, rl, len - l - rl};
            for (int k = 0; k < 4; k++) if (sl[k] > 0) { sub_s[ns] = ss[k]; sub_l[ns] = sl[k]; ns++; }
        }
        int ch[4] = {-1, -1, -1, -1};
        for (int k = 0; k < ns; k++) ch[k] = build(sub_s[k], sub_l[k]);
        Node4& nd = nodes[self];
        nd.leaf = false; nd.nchild = ns;
        for (int k = 0; k < 4; k++) nd.child[k] = ch[k];
        return self;
    }
};
}  // namespace

ORC_API void orc_tree4_gravity(int64_t n, const float* pos, const float* vel, const float* h, const float* m, float dt,
                               float theta, float G, int accum_double, float* grav4, int32_t* num_particles,
                               int32_t* num_approx, int32_t* node_count) {
    std::vector<Aabb> boxes(n);
    std::vector<float> cen(3 * n);
    for (int64_t i = 0; i < n; i++) {
        boxes[i] = ParticleBox(ld3(pos, i), h[i], ld3(vel, i), dt, 0);
        cen[3 * i] = 0.5f * (boxes[i].lo.x + boxes[i].hi.x); cen[3 * i + 1] = 0.5f * (boxes[i].lo.y + boxes[i].hi.y);
        cen[3 * i + 2] = 0.5f * (boxes[i].lo.z + boxes[i].hi.z);
    }
    Tree4 T; T.cen = cen.data(); T.idx.resize(n); std::iota(T.idx.begin(), T.idx.end(), 0);
    T.nodes.reserve(n);
    int root = T.build(0, (int)n);
    // Refit + moments, children before parents (ids of children are larger than their parent's: reverse order)
    for (int k = (int)T.nodes.size() - 1; k >= 0; k--) {
        Node4& nd = T.nodes[k];
        nd.box = {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
        nd.mom = Moment();
        for (int c = 0; c < nd.nchild; c++) {
            if (nd.leaf) { int b = nd.child[c]; nd.box = Union(nd.box, boxes[b]); nd.mom.Accumulate(ld3(pos, b), m[b]); }
            else { const Node4& ch = T.nodes[nd.child[c]]; nd.box = Union(nd.box, ch.box); nd.mom.Accumulate(ch.mom.cm, ch.mom.m); }
        }
    }
    if (node_count) *node_count = (int32_t)T.nodes.size();
#pragma omp parallel
    {
        std::vector<int> stack; stack.reserve(256);
#pragma omp for schedule(dynamic, 64)
        for (int64_t t = 0; t < n; t++) {
            F3 ri = ld3(pos, t); float a = h[t];
            F4 g{0, 0, 0, 0}; double gd[4] = {0, 0, 0, 0}; int np = 0, na = 0;
            auto add = [&](F4 c) { if (accum_double) { gd[0] += c.x; gd[1] += c.y; gd[2] += c.z; gd[3] += c.w; } else { g.x += c.x; g.y += c.y; g.z += c.z; g.w += c.w; } };
            stack.clear(); stack.push_back(root);
            do {
                int k = stack.back(); stack.pop_back();
                const Node4& nd = T.nodes[k];
                if (AcceptApproximation(ri, nd.mom, nd.box, theta)) { add(nd.mom.GravityContribution(ri, G)); na++; }
                else if (nd.leaf) { for (int c = 0; c < nd.nchild; c++) { int b = nd.child[c]; add(GravityContributionParticle(ri, ld3(pos, b), m[b], a, G)); np++; } }
                else for (int c = 0; c < nd.nchild; c++) stack.push_back(nd.child[c]);
            } while (!stack.empty());
            if (accum_double) { g.x = (float)gd[0]; g.y = (float)gd[1]; g.z = (float)gd[2]; g.w = (float)gd[3]; }
            memcpy(grav4 + 4 * t, &g, 16);
            if (num_particles) num_particles[t] = np;
            if (num_approx) num_approx[t] = na;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The reference JOB PATH, stage by stage, with the reference's own parallel / serial structure -- the CPU baseline that
// bench.py times.  Same fp32 arithmetic as the functions above; what differs from
// orc_neighbors_* + orc_density + ... is HOW the pairs are found and stored, i.e. what the reference really pays for:
//   0  ParticleSmoothingSystem (A/Systems/ParticleSmoothingSystem.cs:19-87)                         parallel over bodies
//   1  collider AABBs + 4-ary BVH build + refit (UP/Collision/World/Broadphase.cs:725-782,
//      UP/Collision/Geometry/BoundingVolumeHierarchyBuilder.cs:372-401, 558-600)                     AABBs parallel, build serial
//      (the reference builds the top levels on one thread and <= 64 branches in parallel: a serial build is the conservative stand-in)
//   2  candidate pairs: dual-tree self-overlap of the BVH (Broadphase.cs:275-351,
//      BoundingVolumeHierarchy.cs:107-299): every pair of bodies whose collider AABBs overlap       parallel over tree tasks
//   3  FilterPairs: SplineKernel.Interacts on every candidate (A/Systems/KernelSystem.cs:583-633)   parallel over candidates
//   4  FlattenPairsPrePass (one thread) + FlattenPairs (KernelSystem.cs:539-568, 638-663)            prefix serial, copy parallel
//   5  two stable counting sorts, by body A and by body B, each ONE thread, side by side
//      (ScheduleSortPairsJob / SortPairsSTJob, KernelSystem.cs:339-464)
//   6  CalculateInteractionJob: per body, every pair it is in, BOTH kernels evaluated again on either side (4 kernel
//      evaluations per undirected pair), 40-byte records into the body's buffer (KernelSystem.cs:234-335)   parallel over bodies
//   7  GravityFieldSystem: direct sum, or moments on ONE thread (GenerateMomentsSTJob, GravityFieldSystem.cs:453-555) + the
//      per-particle walk of the 4-ary BVH (:133-215)                                                parallel over bodies
//   8  Integrator.IntegratePosition (UP/Dynamics/Integrator/Integrator.cs:98-101)                  parallel
//   9  DensityFieldSystem (A/Systems/DensityFieldSystem.cs:38-56)                                   parallel over bodies
//  10  PressureFieldSystem: EOS + gradient (A/Systems/PressureFieldSystem.cs:30-70)                 parallel over bodies
//  11  VelocitySystem (A/Systems/VelocitySystem.cs:24-36) + KernelSystem.Cleanup on the main thread (KernelSystem.cs:50-91)
// It is a stand-in for Burst + ECS and FASTER than the real thing: arrays instead of ComponentDataFromEntity lookups, no
// chunk fragmentation, no job-scheduling overhead.
// ------------------------------------------------------------------------------------------------
namespace {
struct BodyPair { int a, b; };
struct IRec { int other; F4 kthis; F4 ksym; };   // ParticleInteraction (A/Components/Kernel.cs:5-16), 36 of its 40 bytes

inline bool Overlap(const Aabb& p, const Aabb& q) {
    return p.lo.x <= q.hi.x && q.lo.x <= p.hi.x && p.lo.y <= q.hi.y && q.lo.y <= p.hi.y && p.lo.z <= q.hi.z && q.lo.z <= p.hi.z;
}
struct OverlapTask { int a, b; };   // b < 0: self-overlap of node a; else node a against node b

struct DualTree {
    const Tree4& T; const std::vector<Aabb>& boxes;
    void emit(std::vector<BodyPair>& out, int x, int y) const { if (Overlap(boxes[x], boxes[y])) out.push_back(x < y ? BodyPair{x, y} : BodyPair{y, x}); }
    void cross(int a, int b, std::vector<BodyPair>& out) const {
        const Node4 &A = T.nodes[a], &B = T.nodes[b];
        if (!Overlap(A.box, B.box)) return;
        if (A.leaf && B.leaf) { for (int c = 0; c < A.nchild; c++) for (int d = 0; d < B.nchild; d++) emit(out, A.child[c], B.child[d]); }
        else if (A.leaf) { for (int d = 0; d < B.nchild; d++) cross(a, B.child[d], out); }
        else if (B.leaf) { for (int c = 0; c < A.nchild; c++) cross(A.child[c], b, out); }
        else for (int c = 0; c < A.nchild; c++) for (int d = 0; d < B.nchild; d++) cross(A.child[c], B.child[d], out);
    }
    void self(int a, std::vector<BodyPair>& out) const {
        const Node4& A = T.nodes[a];
        if (A.leaf) { for (int c = 0; c < A.nchild; c++) for (int d = c + 1; d < A.nchild; d++) emit(out, A.child[c], A.child[d]); return; }
        for (int c = 0; c < A.nchild; c++) self(A.child[c], out);
        for (int c = 0; c < A.nchild; c++) for (int d = c + 1; d < A.nchild; d++) cross(A.child[c], A.child[d], out);
    }
    // one level of a task, as tasks (keeps the emission order of the recursion)
    void expand(const OverlapTask& t, std::vector<OverlapTask>& out) const {
        const Node4& A = T.nodes[t.a];
        if (t.b < 0) {
            if (A.leaf) { out.push_back(t); return; }
            for (int c = 0; c < A.nchild; c++) out.push_back({A.child[c], -1});
            for (int c = 0; c < A.nchild; c++) for (int d = c + 1; d < A.nchild; d++) out.push_back({A.child[c], A.child[d]});
        } else {
            const Node4& B = T.nodes[t.b];
            if (!Overlap(A.box, B.box)) return;
            if (A.leaf || B.leaf) { out.push_back(t); return; }
            for (int c = 0; c < A.nchild; c++) for (int d = 0; d < B.nchild; d++) out.push_back({A.child[c], B.child[d]});
        }
    }
};

inline double now_sec() {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return (double)clock() / CLOCKS_PER_SEC;
#endif
}

void tree4_refit_moments(Tree4& T, const std::vector<Aabb>& boxes, const float* pos, const float* m, bool moments) {
    for (int k = (int)T.nodes.size() - 1; k >= 0; k--) {   // children have larger ids than their parent
        Node4& nd = T.nodes[k];
        if (!moments) nd.box = {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
        else nd.mom = Moment();
        for (int c = 0; c < nd.nchild; c++) {
            if (nd.leaf) { int b = nd.child[c]; if (!moments) nd.box = Union(nd.box, boxes[b]); else nd.mom.Accumulate(ld3(pos, b), m[b]); }
            else { const Node4& ch = T.nodes[nd.child[c]]; if (!moments) nd.box = Union(nd.box, ch.box); else nd.mom.Accumulate(ch.mom.cm, ch.mom.m); }
        }
    }
}
}  // namespace

// gravity: 0 none, 1 direct, 2 tree (4-ary BVH).  pos / vel / h / n_own are updated in place; outputs rho, P, gradP[3n], grav4[4n],
// offsets[n+1] + nbr (the interaction buffers as CSR, in the reference's emission order: pairs where the body is A, then where
// it is B; nbr may be null / nbr_cap too small: only the counts are returned).  counts[0..2] = candidate pairs, filtered
// pairs, directed interactions kept.  stage_sec[12] as listed above.  Returns the number of directed interactions.
ORC_API int64_t orc_reference_step(int64_t n64, float* pos, float* vel, const float* m, float* h, int32_t* n_own, float dt, int gravity,
                                   float K, float G, float theta, float target, int fix_q1, float* rho, float* P, float* gradP,
                                   float* grav4, int32_t* num_particles, int32_t* num_approx, int64_t* offsets, int32_t* nbr,
                                   int64_t nbr_cap, int64_t* counts, double* stage_sec) {
    const int n = (int)n64;
    double t0 = now_sec();
    auto lap = [&](int k) { double t = now_sec(); stage_sec[k] = t - t0; t0 = t; };
    for (int k = 0; k < 12; k++) stage_sec[k] = 0;
    // 0. smoothing lengths from last step's own-support counts
    {
        std::vector<float> hn(n);
        orc_smoothing_update(n, h, n_own, target, hn.data());
        memcpy(h, hn.data(), (size_t)n * 4);
    }
    lap(0);
    // 1. collider AABBs, BVH over their centres, refit
    std::vector<Aabb> boxes(n);
    std::vector<float> cen(3 * (size_t)n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        boxes[i] = ParticleBox(ld3(pos, i), h[i], ld3(vel, i), dt, 0);
        cen[3 * (size_t)i] = 0.5f * (boxes[i].lo.x + boxes[i].hi.x); cen[3 * (size_t)i + 1] = 0.5f * (boxes[i].lo.y + boxes[i].hi.y);
        cen[3 * (size_t)i + 2] = 0.5f * (boxes[i].lo.z + boxes[i].hi.z);
    }
    Tree4 T; T.cen = cen.data(); T.idx.resize(n); std::iota(T.idx.begin(), T.idx.end(), 0);
    T.nodes.reserve(n);
    const int root = n > 0 ? T.build(0, n) : -1;
    tree4_refit_moments(T, boxes, pos, m, false);
    lap(1);
    // 2. candidate pairs: dual-tree self-overlap, parallel over a few thousand sub-tasks
    std::vector<OverlapTask> tasks;
    DualTree D{T, boxes};
    if (root >= 0) tasks.push_back({root, -1});
    for (int round = 0; round < 12 && tasks.size() < 4096; round++) {
        std::vector<OverlapTask> next;
        for (const auto& t : tasks) D.expand(t, next);
        if (next.size() == tasks.size()) break;
        tasks.swap(next);
    }
    std::vector<std::vector<BodyPair>> found(tasks.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = 0; k < (int64_t)tasks.size(); k++) {
        if (tasks[k].b < 0) D.self(tasks[k].a, found[k]); else D.cross(tasks[k].a, tasks[k].b, found[k]);
    }
    std::vector<int64_t> toff(tasks.size() + 1, 0);
    for (size_t k = 0; k < tasks.size(); k++) toff[k + 1] = toff[k] + (int64_t)found[k].size();
    const int64_t ncand = toff[tasks.size()];
    lap(2);
    // 3. FilterPairs: the exact interaction predicate, per candidate (the stream keeps its order)
    std::vector<std::vector<BodyPair>> kept(tasks.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = 0; k < (int64_t)tasks.size(); k++) {
        kept[k].reserve(found[k].size() / 8 + 4);
        for (const BodyPair& p : found[k])
            if (Interacts(ld3(pos, p.a), ld3(pos, p.b), h[p.a], h[p.b])) kept[k].push_back(p);
        std::vector<BodyPair>().swap(found[k]);
    }
    lap(3);
    // 4. flatten: offsets on one thread, copy in parallel
    std::vector<int64_t> koff(tasks.size() + 1, 0);
    for (size_t k = 0; k < tasks.size(); k++) koff[k + 1] = koff[k] + (int64_t)kept[k].size();
    const int64_t npair = koff[tasks.size()];
    std::vector<BodyPair> pairs((size_t)npair);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t k = 0; k < (int64_t)tasks.size(); k++)
        if (!kept[k].empty()) memcpy(&pairs[(size_t)koff[k]], kept[k].data(), kept[k].size() * sizeof(BodyPair));
    kept.clear();
    lap(4);
    // 5. two stable counting sorts (key = body A, key = body B), each on ONE thread, side by side
    std::vector<BodyPair> byA((size_t)npair), byB((size_t)npair);
    std::vector<int64_t> offA((size_t)n + 1, 0), offB((size_t)n + 1, 0);
#pragma omp parallel sections num_threads(2)
    {
#pragma omp section
        {
            for (int64_t e = 0; e < npair; e++) offA[(size_t)pairs[(size_t)e].a + 1]++;
            for (int i = 0; i < n; i++) offA[(size_t)i + 1] += offA[i];
            std::vector<int64_t> cur(offA.begin(), offA.end() - 1);
            for (int64_t e = 0; e < npair; e++) byA[(size_t)cur[pairs[(size_t)e].a]++] = pairs[(size_t)e];
        }
#pragma omp section
        {
            for (int64_t e = 0; e < npair; e++) offB[(size_t)pairs[(size_t)e].b + 1]++;
            for (int i = 0; i < n; i++) offB[(size_t)i + 1] += offB[i];
            std::vector<int64_t> cur(offB.begin(), offB.end() - 1);
            for (int64_t e = 0; e < npair; e++) byB[(size_t)cur[pairs[(size_t)e].b]++] = pairs[(size_t)e];
        }
    }
    lap(5);
    // 6. interaction buffers: per body, both kernels of every pair it is in (the partner evaluates them again)
    std::vector<int64_t> boff((size_t)n + 1, 0);
    for (int i = 0; i < n; i++) boff[(size_t)i + 1] = boff[i] + (offA[(size_t)i + 1] - offA[i]) + (offB[(size_t)i + 1] - offB[i]);
    std::vector<IRec> buf((size_t)boff[n]);
    std::vector<int32_t> bcnt(n, 0);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        IRec* out = buf.data() + boff[i];
        int c = 0;
        const F3 ri = ld3(pos, i);
        for (int64_t e = offA[i]; e < offA[(size_t)i + 1]; e++) {
            const int j = byA[(size_t)e].b;
            Interaction it = CalculateInteraction(ri, ld3(pos, j), h[i], h[j], fix_q1);
            if (it.ksym.w > 0.0f) out[c++] = IRec{j, it.kthis, it.ksym};
        }
        for (int64_t e = offB[i]; e < offB[(size_t)i + 1]; e++) {
            const int j = byB[(size_t)e].a;
            Interaction it = CalculateInteraction(ri, ld3(pos, j), h[i], h[j], fix_q1);
            if (it.ksym.w > 0.0f) out[c++] = IRec{j, it.kthis, it.ksym};
        }
        bcnt[i] = c;
    }
    lap(6);
    // 7. gravity at x_n
    if (gravity == 1) {
        orc_gravity_direct(n, pos, h, m, G, 0, n, 0, grav4);
        if (num_particles) memset(num_particles, 0, (size_t)n * 4);
        if (num_approx) memset(num_approx, 0, (size_t)n * 4);
    } else if (gravity == 2) {
        tree4_refit_moments(T, boxes, pos, m, true);   // GenerateMomentsSTJob: one thread
#pragma omp parallel
        {
            std::vector<int> stack; stack.reserve(256);
#pragma omp for schedule(dynamic, 64)
            for (int t = 0; t < n; t++) {
                const F3 ri = ld3(pos, t); const float a = h[t];
                F4 g{0, 0, 0, 0}; int np = 0, na = 0;
                auto add = [&](F4 c) { g.x += c.x; g.y += c.y; g.z += c.z; g.w += c.w; };
                stack.clear(); stack.push_back(root);
                do {
                    const int k = stack.back(); stack.pop_back();
                    const Node4& nd = T.nodes[k];
                    if (AcceptApproximation(ri, nd.mom, nd.box, theta)) { add(nd.mom.GravityContribution(ri, G)); na++; }
                    else if (nd.leaf) { for (int c = 0; c < nd.nchild; c++) { const int b = nd.child[c]; add(GravityContributionParticle(ri, ld3(pos, b), m[b], a, G)); np++; } }
                    else for (int c = 0; c < nd.nchild; c++) stack.push_back(nd.child[c]);
                } while (!stack.empty());
                memcpy(grav4 + 4 * (size_t)t, &g, 16);
                if (num_particles) num_particles[t] = np;
                if (num_approx) num_approx[t] = na;
            }
        }
    } else {
        memset(grav4, 0, (size_t)n * 16);
        if (num_particles) memset(num_particles, 0, (size_t)n * 4);
        if (num_approx) memset(num_approx, 0, (size_t)n * 4);
    }
    lap(7);
    // 8. x += v dt (old v); the interaction buffers keep the kernels of x_n
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) { pos[3 * (size_t)i] += vel[3 * (size_t)i] * dt; pos[3 * (size_t)i + 1] += vel[3 * (size_t)i + 1] * dt; pos[3 * (size_t)i + 2] += vel[3 * (size_t)i + 2] * dt; }
    lap(8);
    // 9. density + own-support count
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        const IRec* b = buf.data() + boff[i];
        float d = m[i] * Kernel(0.0f, h[i]);
        int own = 0;
        for (int k = 0; k < bcnt[i]; k++) { d += m[b[k].other] * b[k].ksym.w; own += b[k].kthis.w > 0.0f ? 1 : 0; }
        rho[i] = d; n_own[i] = own;
    }
    lap(9);
    // 10. EOS + pressure gradient
    orc_eos(n, rho, K, P);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        const IRec* b = buf.data() + boff[i];
        F3 g{0, 0, 0};
        for (int k = 0; k < bcnt[i]; k++) {
            const int j = b[k].other;
            const float f = m[j] / rho[j] * P[j];
            g.x += f * b[k].ksym.x; g.y += f * b[k].ksym.y; g.z += f * b[k].ksym.z;
        }
        gradP[3 * (size_t)i] = g.x; gradP[3 * (size_t)i + 1] = g.y; gradP[3 * (size_t)i + 2] = g.z;
    }
    lap(10);
    // 11. v += (-gradP/rho - gradPhi) dt; buffers returned / cleared on one thread (KernelSystem.Cleanup)
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            const float a = -gradP[3 * (size_t)i + c] / rho[i] - grav4[4 * (size_t)i + c];
            vel[3 * (size_t)i + c] += a * dt;
        }
    }
    int64_t total = 0;
    offsets[0] = 0;
    for (int i = 0; i < n; i++) {
        if (nbr && total + bcnt[i] <= nbr_cap) for (int k = 0; k < bcnt[i]; k++) nbr[total + k] = buf[(size_t)boff[i] + k].other;
        total += bcnt[i];
        offsets[(size_t)i + 1] = total;
        bcnt[i] = 0;
    }
    lap(11);
    if (counts) { counts[0] = ncand; counts[1] = npair; counts[2] = total; }
    return total;
}
