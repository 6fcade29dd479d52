// kernels_sph.cu -- pressure gradient, near-pair gravity correction, integration, upload/download packing, diagnostics
// (the fused neighbor-list + density + EOS kernel lives in kernels_neighbors.cu).
//
// Replaces: KernelSystem.FilterPairs / CalculateInteractionJob (A/Systems/KernelSystem.cs:234-335, 583-633),
// SplineKernel (A/Util/SplineKernel.cs), DensityFieldSystem (A/Systems/DensityFieldSystem.cs:38-56),
// PressureFieldSystem (A/Systems/PressureFieldSystem.cs:30-70), Integrator.IntegratePosition
// (UP/Dynamics/Integrator/Integrator.cs:98-101) and VelocitySystem (A/Systems/VelocitySystem.cs:24-36).
//
// Numerics contract: *membership* (which j is a neighbor of i, which neighbor is inside i's own support) replays the
// reference's fp32 op sequence exactly (non-contracted __f*_rn intrinsics, IEEE sqrt); *values* (W, grad W, sums) use
// fast arithmetic (reciprocal multiplies, rsqrt, FMA, warp-tree sums) and are held to <= 1e-5 relative by the tests.
#include "ctx.cuh"
#include "rowstream.cuh"
#include <math.h>
#include <string.h>

namespace {

constexpr float kPI = 3.14159274f;          // Mathf.PI
constexpr float kInvPI = 0.318309886f;
constexpr unsigned FULL = 0xffffffffu;

// ---- exact reference arithmetic (SplineKernel.cs:55-89, 115-148), used for edge decisions and the debug surface
__device__ __forceinline__ float kernel_exact(float distance, float size) {
    if (distance >= __fmul_rn(size, 2.0f)) return 0.0f;
    float q = __fdiv_rn(distance, size);
    float pi_h_cube = __fmul_rn(__fmul_rn(__fmul_rn(kPI, size), size), size);
    if (distance < size) {
        float q2 = __fmul_rn(q, q);
        float num = __fadd_rn(__fsub_rn(1.0f, __fmul_rn(1.5f, q2)), __fmul_rn(0.75f, __fmul_rn(q2, q)));
        return __fdiv_rn(num, pi_h_cube);
    }
    float t = __fsub_rn(2.0f, q);
    float num = __fmul_rn(__fmul_rn(t, t), t);
    return __fdiv_rn(num, __fmul_rn(4.0f, pi_h_cube));
}
__device__ __forceinline__ float kernel_deriv_exact(float distance, float size, int fix_q1) {
    if (distance >= __fmul_rn(size, 2.0f)) return 0.0f;
    float q = __fdiv_rn(distance, size);
    float pi_h_4th = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(kPI, size), size), size), size);
    if (distance < size) {
        float q2 = __fmul_rn(q, q);
        float lead = fix_q1 ? -3.0f : 3.0f;
        float num = __fadd_rn(__fmul_rn(lead, q), __fmul_rn(2.25f, q2));
        return __fdiv_rn(num, pi_h_4th);
    }
    float t = __fsub_rn(2.0f, q);
    float num = __fmul_rn(__fmul_rn(-3.0f, t), t);
    return __fdiv_rn(num, __fmul_rn(4.0f, pi_h_4th));
}

// ------------------------------------------------------------------------------------------------------------
// K2: pressure gradient over the materialised lists (PressureFieldSystem.cs:44-70): one warp per 32 consecutive targets,
// rows consumed as a flat stream of octets with canonical summation order (rowstream.cuh).  grad W_sym is recomputed from the
// positions (no stored kernels); the two gathers per pair (posh[j], (m/rho)P of j) are independent and in flight together.
// (dW/dr)/r with the reference's inner branch (quirk Q1: +3q unless lead = -3), c4 = 1/(pi h^4):
//   q < 1 : (lead + 2.25 q) / h * c4 ;   1 <= q < 2 : -0.75 (2-q)^2 / r * c4 ;  0 beyond -- branch-free via t = max(2-q, 0).
// Rows beyond max_neighbors (an error state the caller is told about) contribute their stored part only.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dwr_shape(float q, bool inside, float hinv, float rinv, float lead) {
    const float t = fmaxf(2.0f - q, 0.0f);
    const float outer = (-0.75f * t) * (t * rinv);
    const float inner = fmaf(2.25f, q, lead) * hinv;
    return inside ? inner : outer;
}

constexpr int K2_WARPS = 4;   // 7.6 KB of stream state per warp

__global__ void __launch_bounds__(K2_WARPS * 32, 6) k_pressure_grad(const float4* __restrict__ posh, const float* __restrict__ cvol,
                                                                    const uint32_t* __restrict__ nlist, const int32_t* __restrict__ ncount,
                                                                    int t0, int t1, int rowbase, int kmax, float lead, float4* __restrict__ gradp) {
    __shared__ RowStreamSmem<3> smem[K2_WARPS];   // tg[0]: x, y, z, 1/h   tg[1]: -, h, -, - (read at the spline breakpoint only)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int base = t0 + (blockIdx.x * K2_WARPS + w) * 32;
    if (base >= t1) return;
    RowStreamSmem<3>& S = smem[w];
    const int t = base + lane;
    const bool live = t < t1;
    float4 pi = make_float4(0.f, 0.f, 0.f, 1.f);
    int cnt = 0;
    if (live) { pi = posh[t]; cnt = min(ncount[t], kmax); }
    const float hinv_i = 1.0f / pi.w, h2 = hinv_i * hinv_i;
    S.tg[0][lane] = make_float4(pi.x, pi.y, pi.z, hinv_i);
    S.tg[1][lane] = make_float4(h2 * h2 * kInvPI, pi.w, 0.f, 0.f);
    float acc[3];
    int unused;
    row_stream<3, false>(S, nlist + (size_t)(base - rowbase) * kmax, (uint32_t)base, kmax, cnt,
        [&](const float4 A, const float4* Bp, uint32_t j, float (&v)[3], bool& flag) {
            const float4 pj = posh[j];
            const float cj = cvol[j];
            const float dx = A.x - pj.x, dy = A.y - pj.y, dz = A.z - pj.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            // coincident distinct particles: d = 0, finite factor => gradient contribution 0 (reference: NaN, quirk Q9)
            const float rinv = rs_rsqrt(fmaxf(r2, 1.0e-37f));
            const float r = r2 * rinv;
            const float hinv_j = rs_rcp(pj.w), g2 = hinv_j * hinv_j;
            const float c4_j = g2 * g2 * kInvPI;
            const float qi = r * A.w, qj = r * hinv_j;
            bool in_i = qi < 1.0f, in_j = qj < 1.0f;
            if (fabsf(qi - 1.0f) < 4.0e-6f || fabsf(qj - 1.0f) < 4.0e-6f) {
                // within a few ulps of the spline's breakpoint, where the reference's inner branch (quirk Q1) makes dW/dr jump:
                // decide as the reference does, "distance < size" on the IEEE distance (SplineKernel.cs:115-148)
                const float re = __fsqrt_rn(dot3_rn(__fsub_rn(A.x, pj.x), __fsub_rn(A.y, pj.y), __fsub_rn(A.z, pj.z)));
                in_i = re < Bp->y; in_j = re < pj.w;
            }
            const float a2 = A.w * A.w;
            const float s = fmaf(a2 * a2 * kInvPI, dwr_shape(qi, in_i, A.w, rinv, lead), c4_j * dwr_shape(qj, in_j, hinv_j, rinv, lead)) * cj;
            v[0] = dx * s; v[1] = dy * s; v[2] = dz * s;
            flag = false;
        }, acc, unused);
    if (live) gradp[t] = make_float4(0.5f * acc[0], 0.5f * acc[1], 0.5f * acc[2], 0.f);
}

// Near-pair correction of the gravity kernels, from the neighbor lists: adds, for every listed neighbor j of i,
//     [pair law that should hold]  -  [what the main gravity kernel has already summed for the pair].
// BASE (what was summed): CAPPED = the all-pairs kernel's Newtonian value with r capped at a = h_i (kernels_gravity.cu);
//                         DYER   = the tree walk's P2P law (Dyer & Ip inside a = h_i, Newtonian outside).
// LAW  (what should hold): reference = Dyer & Ip softened inside a = h_i, Newtonian outside (GravityFieldSystem.cs:340-347) --
//     with CAPPED this is non-zero only for r < h_i, and every such pair is in i's list because r < h_i < 2 max(h_i,h_j);
//     SPH_FLAG_PM07_SOFTENING (off by default; roadmap README.md:75-77 "Gravity kernel which conserves energy, see Price &
//     Monaghan 2007") = the force softened with the cubic-spline kernel itself, symmetrised in the two smoothing lengths:
//         grad Phi_i += G m_j (r_i - r_j)/r * 0.5 [phi'(r,h_i) + phi'(r,h_j)],   Phi_i += G m_j 0.5 [phi(r,h_i) + phi(r,h_j)]
//     (P&M 2007 appendix A; phi' = 1/r^2 and phi = -1/r beyond 2h).  It differs from Newton exactly for r < 2 max(h_i,h_j): the
//     neighbor list.  Pairwise antisymmetric (gravity then conserves momentum, unlike the one-sided a = h_i law) and
//     derivable from a potential for fixed h; P&M's extra term for h that varies in time (their zeta / Omega) is not included.
//     In tree mode a listed neighbor that the walk absorbed into an accepted node (possible only under extreme h contrast) is
//     corrected as if it had been summed as a point mass.
__device__ __forceinline__ void pm07_kernel(float r, float hinv, float rinv, float& fr, float& phi) {   // phi'/r and phi of one h
    const float q = r * hinv, q2 = q * q, h3 = hinv * hinv * hinv;
    if (q < 1.0f) {
        fr = h3 * (4.0f / 3.0f + q2 * (-1.2f + 0.5f * q));
        phi = hinv * (q2 * (2.0f / 3.0f + q2 * (-0.3f + 0.1f * q)) - 1.4f);
    } else if (q < 2.0f) {
        const float qi = rinv / hinv;   // 1/q
        fr = h3 * (8.0f / 3.0f - 3.0f * q + q2 * (1.2f - q * (1.0f / 6.0f)) - (1.0f / 15.0f) * qi * qi * qi);
        phi = hinv * (q2 * (4.0f / 3.0f - q + q2 * (0.3f - q * (1.0f / 30.0f))) - 1.6f + (1.0f / 15.0f) * qi);
    } else {
        fr = rinv * rinv * rinv;
        phi = -rinv;
    }
}

constexpr int K2_LPT = 16;

template <bool BASE_CAPPED, bool PM07>
__global__ void __launch_bounds__(256) k_gravity_near(const float4* __restrict__ posh, const float4* __restrict__ posm,
                                                      const uint32_t* __restrict__ nlist, const int32_t* __restrict__ ncount,
                                                      int t0, int t1, int rowbase, int kmax, float G, float4* __restrict__ grav) {
    const int sub = threadIdx.x & (K2_LPT - 1);
    const int t = t0 + (blockIdx.x * blockDim.x + threadIdx.x) / K2_LPT;
    const bool live = t < t1;
    float gx = 0.f, gy = 0.f, gz = 0.f, gp = 0.f;
    if (live) {
        const float4 pi = posh[t];
        const float hinv_i = 1.0f / pi.w;
        const float a2 = pi.w * pi.w;
        const int cnt = min(ncount[t], kmax);
        const uint32_t* row = nlist + (size_t)(t - rowbase) * kmax;
        for (int k = sub; k < cnt; k += K2_LPT) {
            uint32_t j = row[k];
            float4 pj = posm[j];
            float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));   // same expression as the all-pairs kernel
            if (!PM07) {
                if (BASE_CAPPED && r2 < a2) {
                    float r = r2 > 0.f ? r2 * rsqrtf(r2) : 0.f;
                    float x = r * hinv_i, x2 = x * x, x3 = x2 * x;
                    float ma = pj.w * hinv_i;
                    float mg = ma * hinv_i * hinv_i * (7.0f - 9.0f * x + 2.0f * x3);
                    gx = fmaf(dx, mg, gx); gy = fmaf(dy, mg, gy); gz = fmaf(dz, mg, gz);
                    gp -= ma * (1.4f - 4.0f * x2 + 3.0f * x3 - 0.4f * x2 * x3);
                }
            } else {
                const float hj = posh[j].w;
                const float rinv = rsqrtf(fmaxf(r2, 1.0e-37f));
                const float r = r2 * rinv;
                float fi, pi_, fj, pj_;
                pm07_kernel(r, hinv_i, rinv, fi, pi_);
                pm07_kernel(r, 1.0f / hj, rinv, fj, pj_);
                float f = 0.5f * (fi + fj), p = 0.5f * (pi_ + pj_);
                // minus what the main kernel summed
                if (BASE_CAPPED) {
                    const float ci = rsqrtf(fmaxf(r2, a2));
                    f -= ci * ci * ci; p += ci;
                } else if (r2 < a2) {
                    const float x = r * hinv_i, x2 = x * x, x3 = x2 * x;
                    f -= hinv_i * hinv_i * hinv_i * (8.0f - 9.0f * x + 2.0f * x3);
                    p += hinv_i * (2.4f - 4.0f * x2 + 3.0f * x3 - 0.4f * x2 * x3);
                } else {
                    f -= rinv * rinv * rinv; p += rinv;
                }
                const float mg = pj.w * f;
                gx = fmaf(dx, mg, gx); gy = fmaf(dy, mg, gy); gz = fmaf(dz, mg, gz);
                gp = fmaf(pj.w, p, gp);
            }
        }
    }
#pragma unroll
    for (int o = K2_LPT / 2; o > 0; o >>= 1) {
        gx += __shfl_xor_sync(FULL, gx, o); gy += __shfl_xor_sync(FULL, gy, o);
        gz += __shfl_xor_sync(FULL, gz, o); gp += __shfl_xor_sync(FULL, gp, o);
    }
    if (live && sub == 0) {
        float4 g0 = grav[t];
        grav[t] = make_float4(g0.x + G * gx, g0.y + G * gy, g0.z + G * gz, g0.w + G * gp);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Integrate: x += v*dt (old v), v += (-gradP/rho - gradPhi)*dt ; exact op order of the reference.
// kick_drift (SPH_FLAG_KICK_DRIFT, off by default; roadmap README.md:90-93): v first, then x += v_new*dt -- the
// symplectic (leapfrog with staggered velocities) variant of the same one-force-evaluation step.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_integrate(float4* __restrict__ posh, float4* __restrict__ velm,
                                                   const float* __restrict__ rho, const float4* __restrict__ gradp,
                                                   const float4* __restrict__ grav, int t0, int t1, float dt, int kick_drift) {
    int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t1) return;
    float4 p = posh[t], v = velm[t], gp = gradp[t], g = grav[t];
    float d = rho[t];
    if (!kick_drift) {
        p.x = __fadd_rn(p.x, __fmul_rn(v.x, dt));
        p.y = __fadd_rn(p.y, __fmul_rn(v.y, dt));
        p.z = __fadd_rn(p.z, __fmul_rn(v.z, dt));
    }
    float ax = __fsub_rn(__fdiv_rn(-gp.x, d), g.x);
    float ay = __fsub_rn(__fdiv_rn(-gp.y, d), g.y);
    float az = __fsub_rn(__fdiv_rn(-gp.z, d), g.z);
    v.x = __fadd_rn(v.x, __fmul_rn(ax, dt));
    v.y = __fadd_rn(v.y, __fmul_rn(ay, dt));
    v.z = __fadd_rn(v.z, __fmul_rn(az, dt));
    if (kick_drift) {
        p.x = __fadd_rn(p.x, __fmul_rn(v.x, dt));
        p.y = __fadd_rn(p.y, __fmul_rn(v.y, dt));
        p.z = __fadd_rn(p.z, __fmul_rn(v.z, dt));
    }
    posh[t] = p;
    velm[t] = v;
}

// ------------------------------------------------------------------------------------------------------------
// Upload / download packing
// ------------------------------------------------------------------------------------------------------------
// staging (device): pos[3n] vel[3n] mass[n] h[n] nown[n]
// has_nown: 0 = h only (st[7n..8n)), 1 = h[n] then nown[n], 2 = raw sph_ParticleSmoothing records (7 words each) at st[7n..14n)
// Also reduces the range of the masses into mm[0..1] (ordered uints): equal masses select the hoisted-mass kernels.
__global__ void __launch_bounds__(256) k_pack_upload(const float* __restrict__ st, int n, int has_nown, uint32_t orig0, float4* __restrict__ posh,
                                                     float4* __restrict__ velm, uint32_t* __restrict__ orig,
                                                     int32_t* __restrict__ nown, uint32_t* __restrict__ mm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    {
        const uint32_t mo = i < n ? f2ord(st[6 * (size_t)n + i]) : 0u;
        const uint32_t lo = __reduce_min_sync(FULL, i < n ? mo : 0xffffffffu), hi = __reduce_max_sync(FULL, mo);
        if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
        // largest uploaded h (NaN and negative values count as +inf: the host then keeps the literal-kernel pass armed)
        float hv = 0.f;
        if (i < n) { hv = has_nown == 2 ? st[7 * (size_t)n + 7 * (size_t)i] : st[7 * (size_t)n + i]; if (!(hv >= 0.f)) hv = INFINITY; }
        const uint32_t hm = __reduce_max_sync(FULL, __float_as_uint(hv));
        if ((threadIdx.x & 31) == 0) atomicMax(&mm[2], hm);
    }
    if (i >= n) return;
    const float* pos = st;
    const float* vel = st + 3 * (size_t)n;
    const float* mass = st + 6 * (size_t)n;
    const float* h = st + 7 * (size_t)n;
    const int32_t* no = (const int32_t*)(st + 8 * (size_t)n);
    const float hi = has_nown == 2 ? h[7 * (size_t)i] : h[i];
    posh[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], hi);
    velm[i] = make_float4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], mass[i]);
    orig[i] = orig0 + (uint32_t)i;
    nown[i] = has_nown == 2 ? ((const int32_t*)h)[7 * (size_t)i + 6] : has_nown ? no[i] : 0;
}

__global__ void __launch_bounds__(256) k_unpack_field(int field, int n, const uint32_t* __restrict__ orig,
                                                      const float4* __restrict__ posh, const float4* __restrict__ velm,
                                                      const float* __restrict__ rho, const float* __restrict__ press,
                                                      const float4* __restrict__ gradp, const float4* __restrict__ grav,
                                                      const int32_t* __restrict__ npart, const int32_t* __restrict__ napprox,
                                                      const int32_t* __restrict__ ncount, const int32_t* __restrict__ nown,
                                                      float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t b = orig[i];
    switch (field) {
        case SPH_FIELD_TRANSLATION: { float4 p = posh[i]; out[3 * b] = p.x; out[3 * b + 1] = p.y; out[3 * b + 2] = p.z; break; }
        case SPH_FIELD_VELOCITY: { float4 v = velm[i]; out[3 * b] = v.x; out[3 * b + 1] = v.y; out[3 * b + 2] = v.z; break; }
        case SPH_FIELD_MASS: out[b] = velm[i].w; break;
        case SPH_FIELD_SMOOTHING: out[2 * b] = posh[i].w; ((int32_t*)out)[2 * b + 1] = nown[i]; break;
        case SPH_FIELD_COUNT_: {   // the full sph_ParticleSmoothing record (ParticleSmoothing.cs:16-31), 7 words
            const float h = posh[i].w;
            float* o = out + 7 * b;
            o[0] = h; o[1] = 2.0f * h; o[2] = 0.f; o[3] = 0.f; o[4] = 0.f; o[5] = 2.0f * h;
            ((int32_t*)o)[6] = nown[i];
            break;
        }
        case SPH_FIELD_DENSITY: out[b] = rho[i]; break;
        case SPH_FIELD_PRESSURE: out[b] = press[i]; break;
        case SPH_FIELD_PRESSURE_GRAD: { float4 q = gradp[i]; out[3 * b] = q.x; out[3 * b + 1] = q.y; out[3 * b + 2] = q.z; break; }
        case SPH_FIELD_GRAVITY: {
            float4 q = grav[i];
            out[6 * b] = q.x; out[6 * b + 1] = q.y; out[6 * b + 2] = q.z; out[6 * b + 3] = q.w;
            ((int32_t*)out)[6 * b + 4] = npart[i]; ((int32_t*)out)[6 * b + 5] = napprox[i];
            break;
        }
        case SPH_FIELD_NEIGHBOR_COUNT: ((int32_t*)out)[b] = ncount[i]; break;
    }
}

// One warp per sorted slot: map the list row to body indices, bitonic-sort ascending, write at offsets[body].
__global__ void __launch_bounds__(128) k_neighbor_rows(const uint32_t* __restrict__ nlist, const int32_t* __restrict__ ncount,
                                                       const uint32_t* __restrict__ orig, const int64_t* __restrict__ offsets,
                                                       int n, int rowbase, int kmax, int p2, int32_t* __restrict__ nbr) {
    extern __shared__ uint32_t sm[];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int t = blockIdx.x * 4 + w;
    if (t >= n) return;
    uint32_t* a = sm + (size_t)w * p2;
    int cnt = min(ncount[t], kmax);
    for (int k = lane; k < p2; k += 32) a[k] = k < cnt ? orig[nlist[(size_t)(t - rowbase) * kmax + k]] : 0xffffffffu;
    __syncwarp();
    for (int size = 2; size <= p2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int k = lane; k < p2; k += 32) {
                int partner = k ^ stride;
                if (partner > k) {
                    bool up = (k & size) == 0;
                    uint32_t x = a[k], y = a[partner];
                    if ((x > y) == up) { a[k] = y; a[partner] = x; }
                }
            }
            __syncwarp();
        }
    int64_t o = offsets[orig[t]];
    for (int k = lane; k < cnt; k += 32) nbr[o + k] = (int32_t)a[k];
}

__global__ void __launch_bounds__(256) k_inverse_map(const uint32_t* __restrict__ orig, int n, uint32_t* __restrict__ inv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[orig[i]] = (uint32_t)i;
}

// Debug/parity surface: the reference's interaction record, computed with the exact op sequence
// (KernelSystem.cs:305-334, SplineKernel.cs:102-111).  One thread per body.
__global__ void __launch_bounds__(128) k_interactions(const float4* __restrict__ posh, const uint32_t* __restrict__ inv, int n,
                                                      const int64_t* __restrict__ offsets, const int32_t* __restrict__ nbr,
                                                      int fix_q1, sph_ParticleInteraction* __restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    float4 pi = posh[inv[b]];
    for (int64_t e = offsets[b]; e < offsets[b + 1]; e++) {
        int jb = nbr[e];
        float4 pj = posh[inv[jb]];
        float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y), dz = __fsub_rn(pi.z, pj.z);
        float dist = __fsqrt_rn(dot3_rn(dx, dy, dz));
        float si = __fdiv_rn(kernel_deriv_exact(dist, pi.w, fix_q1), dist);
        float sj = __fdiv_rn(kernel_deriv_exact(dist, pj.w, fix_q1), dist);
        float wi = kernel_exact(dist, pi.w), wj = kernel_exact(dist, pj.w);
        float ki[4] = {__fmul_rn(dx, si), __fmul_rn(dy, si), __fmul_rn(dz, si), wi};
        float kj[4] = {__fmul_rn(dx, sj), __fmul_rn(dy, sj), __fmul_rn(dz, sj), wj};
        sph_ParticleInteraction it;
        it.otherIndex = jb; it.otherVersion = 1;
        for (int k = 0; k < 4; k++) { it.kernelThis[k] = ki[k]; it.kernelSymmetric[k] = __fmul_rn(__fadd_rn(ki[k], kj[k]), 0.5f); }
        out[e] = it;
    }
}

// Diagnostics: block reduce in double, one atomicAdd per block per quantity.
__global__ void __launch_bounds__(256) k_diagnostics(const float4* __restrict__ posh, const float4* __restrict__ velm,
                                                     const float* __restrict__ rho, const float4* __restrict__ grav,
                                                     const int32_t* __restrict__ ncount, int n, float Keos, double* __restrict__ out,
                                                     int* __restrict__ maxcount) {
    double v[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int mc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = posh[i], u = velm[i];
        double m = u.w;
        v[0] += m;
        v[1] += m * u.x; v[2] += m * u.y; v[3] += m * u.z;
        v[4] += m * ((double)p.y * u.z - (double)p.z * u.y);
        v[5] += m * ((double)p.z * u.x - (double)p.x * u.z);
        v[6] += m * ((double)p.x * u.y - (double)p.y * u.x);
        v[7] += 0.5 * m * ((double)u.x * u.x + (double)u.y * u.y + (double)u.z * u.z);
        v[8] += 0.5 * m * (double)grav[i].w;
        v[9] += m * (double)Keos * (double)rho[i];
        v[10] += (double)ncount[i];
        mc = max(mc, ncount[i]);
    }
    __shared__ double s[11][8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < 11; k++) {
        double x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
        if (lane == 0) s[k][w] = x;
    }
    for (int o = 16; o > 0; o >>= 1) mc = max(mc, __shfl_xor_sync(FULL, mc, o));
    if (lane == 0) atomicMax(maxcount, mc);
    __syncthreads();
    if (threadIdx.x < 11) {
        double x = 0;
        for (int j = 0; j < 8; j++) x += s[threadIdx.x][j];
        atomicAdd(&out[threadIdx.x], x);
    }
}

// Field statistics (README.md:50-52 roadmap: "Average/max/min: Temp, Pressure, Density, Grav Field"): sum (fp64), min and max
// (ordered-uint atomics) of rho, P, |grad Phi| and u = K rho (the specific internal energy of the P = K rho^2 gas: the model's
// temperature proxy).  out[0..3] sums; mm[0..3] ordered minima, mm[4..7] ordered maxima.
__global__ void __launch_bounds__(256) k_field_stats(const float* __restrict__ rho, const float* __restrict__ press,
                                                     const float4* __restrict__ grav, int n, float Keos, double* __restrict__ out,
                                                     uint32_t* __restrict__ mm) {
    double s[4] = {0, 0, 0, 0};
    float lo[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, hi[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 g = grav[i];
        const float d = rho[i];
        const float q[4] = {d, press[i], sqrtf(g.x * g.x + g.y * g.y + g.z * g.z), Keos * d};
#pragma unroll
        for (int k = 0; k < 4; k++) { s[k] += (double)q[k]; lo[k] = fminf(lo[k], q[k]); hi[k] = fmaxf(hi[k], q[k]); }
    }
    __shared__ double ss[4][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double x = s[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
        const uint32_t l = __reduce_min_sync(FULL, f2ord(lo[k])), h = __reduce_max_sync(FULL, f2ord(hi[k]));
        if (lane == 0) { ss[k][w] = x; atomicMin(&mm[k], l); atomicMax(&mm[4 + k], h); }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double x = 0;
        for (int j = 0; j < 8; j++) x += ss[threadIdx.x][j];
        atomicAdd(&out[threadIdx.x], x);
    }
}

}  // namespace

// ---- launchers ------------------------------------------------------------------------------------------------
static inline void target_range(sphb200_ctx* c, int& t0, int& t1) {
    t0 = (int)c->t0;
    t1 = (c->t1 < 0 || c->t1 > c->n) ? (int)c->n : (int)c->t1;
    if (t0 > t1) t0 = t1;
}

int sph_launch_pressure(sphb200_ctx* c) {
    int t0, t1; target_range(c, t0, t1);
    int nt = t1 - t0;
    if (nt <= 0) return SPH_OK;
    float lead = (c->p.flags & SPH_FLAG_FIX_KERNEL_DERIV_SIGN) ? -3.0f : 3.0f;
    int tpb = K2_WARPS * 32;   // targets per block: 32 per warp
    k_pressure_grad<<<sph_div_up(nt, tpb), K2_WARPS * 32, 0, c->stream>>>(c->posh[c->cur], c->cvol, c->nlist, c->ncount, t0, t1,
                                                                (int)c->row_base, c->p.max_neighbors, lead, c->gradp);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

// impl = the gravity kernel that has just run (SPH_GRAVITY_PARTICLE / SPH_GRAVITY_TREE); see k_gravity_near
int sph_launch_gravity_near(sphb200_ctx* c, int impl) {
    int t0, t1; target_range(c, t0, t1);
    int nt = t1 - t0;
    if (nt <= 0) return SPH_OK;
    const bool pm07 = (c->p.flags & SPH_FLAG_PM07_SOFTENING) != 0;
    if (impl == SPH_GRAVITY_TREE && !pm07) return SPH_OK;      // the walk applies the reference's softened law itself
    int tpb = 256 / K2_LPT;
#define NEAR_LAUNCH(B, P) k_gravity_near<B, P><<<sph_div_up(nt, tpb), 256, 0, c->stream>>>(c->posh[c->cur], c->posm, c->nlist, c->ncount, t0, t1, \
                                                                                        (int)c->row_base, c->p.max_neighbors, c->p.G, c->grav)
    if (impl == SPH_GRAVITY_PARTICLE) { if (pm07) NEAR_LAUNCH(true, true); else NEAR_LAUNCH(true, false); }
    else NEAR_LAUNCH(false, true);
#undef NEAR_LAUNCH
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

int sph_launch_integrate(sphb200_ctx* c, float dt) {
    int t0, t1; target_range(c, t0, t1);
    int nt = t1 - t0;
    if (nt <= 0) return SPH_OK;
    k_integrate<<<sph_div_up(nt, 256), 256, 0, c->stream>>>(c->posh[c->cur], c->velm[c->cur], c->rho, c->gradp, c->grav, t0, t1, dt,
                                                          (c->p.flags & SPH_FLAG_KICK_DRIFT) ? 1 : 0);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

int sph_launch_pack_upload(sphb200_ctx* c, int64_t n, int has_nown, uint32_t orig0) {
    k_pack_upload<<<sph_div_up(n, 256), 256, 0, c->stream>>>((const float*)c->stage_d, (int)n, has_nown, orig0, c->posh[0],
                                                            c->velm[0], c->orig[0], c->nown, c->bounds + 12);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

int sph_launch_unpack_field(sphb200_ctx* c, int field, int* elem_bytes) {
    static const int words[SPH_FIELD_COUNT_ + 1] = {3, 3, 1, 2, 1, 1, 3, 6, 1, 7};   // [SPH_FIELD_COUNT_] = full smoothing record
    if (field < 0 || field > SPH_FIELD_COUNT_) return SPH_ERR_INVALID_ARG;
    *elem_bytes = words[field] * 4;
    int n = (int)c->n;
    k_unpack_field<<<sph_div_up(n, 256), 256, 0, c->stream>>>(field, n, c->orig[c->cur], c->posh[c->cur], c->velm[c->cur], c->rho,
                                                             c->press, c->gradp, c->grav, c->npart, c->napprox, c->ncount,
                                                             c->nown, (float*)c->stage_d);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

// offsets_d: int64[n+1] in body order (device); rows_d: int32[total]
int sph_launch_neighbor_rows_sorted(sphb200_ctx* c, int32_t* rows_d) {
    int n = (int)c->n;
    int kmax = c->p.max_neighbors;
    int p2 = 32;
    while (p2 < kmax) p2 <<= 1;
    const int64_t* offsets_d = (const int64_t*)c->stage_d;
    k_neighbor_rows<<<sph_div_up(n, 4), 128, 4 * p2 * sizeof(uint32_t), c->stream>>>(c->nlist, c->ncount, c->orig[c->cur], offsets_d, n,
                                                                                    (int)c->row_base, kmax, p2, rows_d);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

int sph_launch_interactions(sphb200_ctx* c, int64_t total, const int64_t* offsets_d, const int32_t* nbr_d,
                            sph_ParticleInteraction* out_d) {
    (void)total;
    int n = (int)c->n;
    uint32_t* inv = c->idx[0];  // scratch (free outside build_neighbors)
    k_inverse_map<<<sph_div_up(n, 256), 256, 0, c->stream>>>(c->orig[c->cur], n, inv);
    SPH_LAUNCH_CHECK(c);
    k_interactions<<<sph_div_up(n, 128), 128, 0, c->stream>>>(c->posh[c->cur], inv, n, offsets_d, nbr_d,
                                                             (c->p.flags & SPH_FLAG_FIX_KERNEL_DERIV_SIGN) ? 1 : 0, out_d);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

// sums over resident slots [off, off+n) into diag_d[0..10], max neighbor count into the int at diag_d[12]
int sph_launch_diagnostics_range(sphb200_ctx* c, int off, int n) {
    SPH_CK(c, cudaMemsetAsync(c->diag_d, 0, 16 * sizeof(double), c->stream));
    if (n <= 0) return SPH_OK;
    int* maxc = (int*)(c->diag_d + 12);
    int blocks = min(sph_div_up(n, 256), c->sm_count * 4);
    k_diagnostics<<<blocks, 256, 0, c->stream>>>(c->posh[c->cur] + off, c->velm[c->cur] + off, c->rho + off, c->grav + off, c->ncount + off, n,
                                                 c->p.K, c->diag_d, maxc);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

int sph_launch_diagnostics(sphb200_ctx* c, double* out12) {
    int n = (int)c->n;
    int rc = sph_launch_diagnostics_range(c, 0, n);
    if (rc) return rc;
    double tmp[16];
    SPH_CK(c, cudaMemcpyAsync(tmp, c->diag_d, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 10; k++) out12[k] = tmp[k];
    out12[10] = n > 0 ? tmp[10] / n : 0.0;
    out12[11] = (double)(*(int*)&tmp[12]);
    return SPH_OK;
}

// sums of rho, P, |grad Phi|, K rho over resident slots [off, off+n) into diag_d[16..19]; ordered-uint minima / maxima into the
// eight uint32 at diag_d + 20
int sph_launch_field_stats_range(sphb200_ctx* c, int off, int n) {
    uint32_t* mm = (uint32_t*)(c->diag_d + 20);
    SPH_CK(c, cudaMemsetAsync(c->diag_d + 16, 0, 4 * sizeof(double), c->stream));
    SPH_CK(c, cudaMemsetAsync(mm, 0xff, 4 * sizeof(uint32_t), c->stream));
    SPH_CK(c, cudaMemsetAsync(mm + 4, 0, 4 * sizeof(uint32_t), c->stream));
    if (n <= 0) return SPH_OK;
    int blocks = min(sph_div_up(n, 256), c->sm_count * 4);
    k_field_stats<<<blocks, 256, 0, c->stream>>>(c->rho + off, c->press + off, c->grav + off, n, c->p.K, c->diag_d + 16, mm);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

// out12 = (min, max, mean) x (rho, P, |grad Phi|, K rho) from the sums / ordered extrema (host side; n = particles they cover)
void sph_finish_field_stats(const double* sums4, const uint32_t* mm8, int64_t n, double* out12) {
    auto ord2f_h = [](uint32_t u) { uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &b, 4); return (double)f; };
    for (int k = 0; k < 4; k++) {
        out12[3 * k] = n > 0 ? ord2f_h(mm8[k]) : 0.0;
        out12[3 * k + 1] = n > 0 ? ord2f_h(mm8[4 + k]) : 0.0;
        out12[3 * k + 2] = n > 0 ? sums4[k] / (double)n : 0.0;
    }
}
