// kernels_tree.cu -- Barnes-Hut tree gravity on an LBVH built over the Morton-sorted particles.
//
// Replaces: the Unity.Physics 4-ary broadphase BVH used as a Barnes-Hut tree
// (UP/Collision/Geometry/BoundingVolumeHierarchy.cs:40-83, ...Builder.cs:416-467), GenerateMomentsSTJob
// (A/Systems/GravityFieldSystem.cs:453-555), the per-particle tree walk (:133-215), the Bmax MAC (:229-247) and
// the moment arithmetic (:367-443).  Tree *shape* is not a parity target; what is kept exactly is
// the reference's moment/MAC arithmetic and walk semantics, executed on this LBVH.  oracle/sph_oracle.cpp
// (orc_lbvh_topology / orc_lbvh_moments / orc_tree_walk) builds the identical tree on the CPU.
//
// Node ids: internal 0..n-2 (root 0), leaf of sorted slot s -> n-1+s.  A node holding <= leaf_max particles plays
// the role of a reference leaf node (never descended; bodies summed directly, including the target itself, Q3).
//
// Walk mapping: one warp = 32 consecutive (spatially coherent) sorted targets sharing ONE traversal stack in shared
// memory; each stack entry carries the mask of lanes still descending that subtree, so every lane sees exactly the
// node sequence of its private depth-first walk (per-particle MAC, same summation order as the oracle) while node
// loads are warp-uniform broadcasts and control flow is warp-coherent.  The walk reads a packed 32-byte node
// (centre of mass, mass, Bmax^2, children / body range): Bmax^2 depends only on the node, so it is evaluated once in
// the build with the reference's exact op sequence instead of once per (target, node) visit.
#include "ctx.cuh"
#include <math.h>

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ int delta_fn(const uint32_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    return a == b ? 32 + __clz((uint32_t)i ^ (uint32_t)j) : __clz(a ^ b);
}

// Karras (2012) binary radix tree over (key, slot) -- integer-only, identical to orc_lbvh_topology.
__global__ void __launch_bounds__(256) k_lbvh_topology(const uint32_t* __restrict__ keys, int n, int2* __restrict__ child,
                                                       int2* __restrict__ range, int32_t* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    child[n - 1 + i] = make_int2(-1, -1);
    range[n - 1 + i] = make_int2(i, i);
    if (i == 0) parent[0] = -1;
    if (i >= n - 1) return;
    int d = (delta_fn(keys, n, i, i + 1) - delta_fn(keys, n, i, i - 1)) > 0 ? 1 : -1;
    int dmin = delta_fn(keys, n, i, i - d);
    int lmax = 2;
    while (delta_fn(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta_fn(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta_fn(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta_fn(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
    int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    child[i] = make_int2(lc, rc);
    range[i] = make_int2(lo, hi);
    parent[lc] = i;
    parent[rc] = i;
}

// GravitationalMoment.Accumulate (GravityFieldSystem.cs:398-411), exact op order
__device__ __forceinline__ void moment_accumulate(float4& mo, float cx, float cy, float cz, float first) {
    if (first != 0.0f) {
        float nm = __fadd_rn(mo.w, first);
        mo.x = __fdiv_rn(__fadd_rn(__fmul_rn(mo.x, mo.w), __fmul_rn(cx, first)), nm);
        mo.y = __fdiv_rn(__fadd_rn(__fmul_rn(mo.y, mo.w), __fmul_rn(cy, first)), nm);
        mo.z = __fdiv_rn(__fadd_rn(__fmul_rn(mo.z, mo.w), __fmul_rn(cz, first)), nm);
        mo.w = nm;
    }
}

// Packed walk node: [2k] = (cm.xyz, M), [2k+1] = (Bmax^2, a, b, -) with (a,b) = (left,right) for nodes that are
// descended and (first, -count) for leaf buckets.  Bmax^2 with the op order of AcceptApproximation (:235-243).
__device__ __forceinline__ void pack_node(float4* __restrict__ packed, int k, float4 mo, const float lo[3], const float hi[3],
                                          int2 ch, int2 rg, int leaf_max) {
    float bx = fmaxf(__fsub_rn(hi[0], mo.x), __fsub_rn(mo.x, lo[0]));
    float by = fmaxf(__fsub_rn(hi[1], mo.y), __fsub_rn(mo.y, lo[1]));
    float bz = fmaxf(__fsub_rn(hi[2], mo.z), __fsub_rn(mo.z, lo[2]));
    float b_sq = dot3_rn(bx, by, bz);
    int cnt = rg.y - rg.x + 1;
    int a = cnt <= leaf_max ? rg.x : ch.x;
    int b = cnt <= leaf_max ? -cnt : ch.y;
    packed[2 * (size_t)k] = mo;
    packed[2 * (size_t)k + 1] = make_float4(b_sq, __int_as_float(a), __int_as_float(b), 0.f);
}

// Moments + MAC boxes.  Small nodes (<= leaf_max bodies) are evaluated directly from their particle range (reference
// leaf rule); larger nodes are finished bottom-up by the second thread to arrive (fixed left-then-right order).
__global__ void __launch_bounds__(256) k_lbvh_nodes(const float4* __restrict__ posh, const float4* __restrict__ velm, int n,
                                                    const int2* __restrict__ child, const int2* __restrict__ range,
                                                    const int32_t* __restrict__ parent, int leaf_max, int aabb_mode, float dt,
                                                    int32_t* __restrict__ flag, float4* mom, float4* nlo, float4* nhi,
                                                    float4* __restrict__ packed) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * n - 1) return;
    int2 rg = range[k];
    if (rg.y - rg.x + 1 > leaf_max) return;
    float4 mo = make_float4(0.f, 0.f, 0.f, 0.f);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const float margin = 0.1f * 0.5f;  // CollisionTolerance * 0.5 (Broadphase.cs:200, CollisionWorld.cs:32)
    for (int s = rg.x; s <= rg.y; s++) {
        float4 p = posh[s], v = velm[s];
        moment_accumulate(mo, p.x, p.y, p.z, v.w);
        float x[3] = {p.x, p.y, p.z}, vel[3] = {v.x, v.y, v.z};
        for (int c = 0; c < 3; c++) {
            float bl, bh;
            if (aabb_mode == 1) { bl = x[c]; bh = x[c]; }
            else {
                // sphere AABB radius 2h (Physics_SphereCollider.cs:129-137), swept by v*dt (Motion.cs:142-146), +-margin
                float rad = __fmul_rn(p.w, 2.0f);
                bl = __fsub_rn(x[c], rad); bh = __fadd_rn(x[c], rad);
                float lin = __fmul_rn(vel[c], dt);
                bh = __fadd_rn(fmaxf(bh, __fadd_rn(bh, lin)), 0.0f);
                bl = __fsub_rn(fminf(bl, __fadd_rn(bl, lin)), 0.0f);
                bl = __fsub_rn(bl, margin); bh = __fadd_rn(bh, margin);
            }
            lo[c] = fminf(lo[c], bl); hi[c] = fmaxf(hi[c], bh);
        }
    }
    mom[k] = mo;
    nlo[k] = make_float4(lo[0], lo[1], lo[2], __int_as_float(rg.x));
    nhi[k] = make_float4(hi[0], hi[1], hi[2], __int_as_float(rg.y));
    pack_node(packed, k, mo, lo, hi, child[k], rg, leaf_max);
    int cur = k;
    while (true) {
        int p = parent[cur];
        if (p < 0) break;
        int2 prg = range[p];
        if (prg.y - prg.x + 1 <= leaf_max) break;  // parent is a small node, evaluated by its own thread
        __threadfence();
        if (atomicAdd(&flag[p], 1) == 0) break;    // first arrival: sibling not ready yet
        int2 ch = child[p];
        float4 ml = __ldcg(&mom[ch.x]), mr = __ldcg(&mom[ch.y]);
        float4 ll = __ldcg(&nlo[ch.x]), lr = __ldcg(&nlo[ch.y]);
        float4 hl = __ldcg(&nhi[ch.x]), hr = __ldcg(&nhi[ch.y]);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        moment_accumulate(acc, ml.x, ml.y, ml.z, ml.w);  // GravityFieldSystem.cs:513-521, children in Data order
        moment_accumulate(acc, mr.x, mr.y, mr.z, mr.w);
        float plo[3] = {fminf(ll.x, lr.x), fminf(ll.y, lr.y), fminf(ll.z, lr.z)};
        float phi[3] = {fmaxf(hl.x, hr.x), fmaxf(hl.y, hr.y), fmaxf(hl.z, hr.z)};
        mom[p] = acc;
        nlo[p] = make_float4(plo[0], plo[1], plo[2], __int_as_float(prg.x));
        nhi[p] = make_float4(phi[0], phi[1], phi[2], __int_as_float(prg.y));
        pack_node(packed, p, acc, plo, phi, ch, prg, leaf_max);
        cur = p;
    }
}

constexpr int TW_WARPS = 8;

struct WalkAcc {
    float gx, gy, gz, gp;
    int np, na;
};

// AcceptApproximation (GravityFieldSystem.cs:229-247): bmax_sq / r_sq < theta^2 with the exact op order for r_sq; the
// quotient is taken with a fast reciprocal and re-done with the IEEE division only inside a band around theta^2.
__device__ __forceinline__ bool mac_accept(const float4& pi, const float4& A, float b_sq, float theta2, float band, float& dx,
                                           float& dy, float& dz, float& r_sq) {
    dx = __fsub_rn(pi.x, A.x); dy = __fsub_rn(pi.y, A.y); dz = __fsub_rn(pi.z, A.z);
    r_sq = dot3_rn(dx, dy, dz);
    const float q = b_sq * rcp_approx(r_sq);
    bool acc = q < theta2;
    if (fabsf(q - theta2) < band) acc = __fdiv_rn(b_sq, r_sq) < theta2;   // rare: the IEEE quotient decides
    return acc;
}

// GravitationalMoment.GravityContribution (M2P, :428-442)
__device__ __forceinline__ void m2p(WalkAcc& w, const float4& A, float dx, float dy, float dz, float r_sq) {
    float rinv = rsqrt_approx(r_sq);
    float mr = A.w * rinv;
    float g = mr * rinv * rinv;
    w.gx = fmaf(dx, g, w.gx); w.gy = fmaf(dy, g, w.gy); w.gz = fmaf(dz, g, w.gz);
    w.gp -= mr;
    w.na++;
}

// Leaf bucket: bodies first .. first+count-1 summed directly with GravityContributionParticle (:332-356), a = h_i;
// includes the target itself (quirk Q3).  `soft`: some open lane may be inside its softening radius.
__device__ __forceinline__ void p2p_bucket(WalkAcc& w, const float4& pi, float a2, float ainv, const float4* __restrict__ posm,
                                           int first, int cnt, bool open, bool soft) {
    for (int s = first; s < first + cnt; s++) {
        const float4 pj = __ldg(&posm[s]);
        float ex = pi.x - pj.x, ey = pi.y - pj.y, ez = pi.z - pj.z;
        float r2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
        float rinv = rsqrt_approx(fmaxf(r2, a2));
        float mr = pj.w * rinv;
        float g = mr * rinv * rinv, ph = -mr;
        if (soft && r2 < a2) {
            float r = r2 > 0.f ? r2 * rsqrt_approx(r2) : 0.f;
            float x = r * ainv, x2 = x * x, x3 = x2 * x;
            float ma = pj.w * ainv;
            g = ma * ainv * ainv * (8.0f - 9.0f * x + 2.0f * x3);
            ph = -ma * (2.4f - 4.0f * x2 + 3.0f * x3 - 0.4f * x2 * x3);
        }
        if (open) {
            w.gx = fmaf(ex, g, w.gx); w.gy = fmaf(ey, g, w.gy); w.gz = fmaf(ez, g, w.gz);
            w.gp += ph;
            w.np++;
        }
    }
}

// One traversal step handles BOTH children of an opened node (4 independent node loads in flight, two MAC tests
// interleaved); a stack entry is (left, right, mask of lanes that rejected the parent).  Every lane still sees exactly
// the accepted nodes / opened buckets of its private walk (per-particle MAC); only the order in which a lane adds its
// contributions differs from the depth-first order of the oracle.
__global__ void __launch_bounds__(TW_WARPS * 32) k_tree_walk(const float4* __restrict__ posh, const float4* __restrict__ posm,
                                                             const float4* __restrict__ packed, int t0, int t1, float theta2,
                                                             float G, float4* __restrict__ grav, int32_t* __restrict__ npart,
                                                             int32_t* __restrict__ napprox, int32_t* __restrict__ err) {
    __shared__ int4 stack[TW_WARPS][SPH_TREE_STACK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t = t0 + (blockIdx.x * TW_WARPS + wid) * 32 + lane;
    const bool active = t < t1;
    const float4 pi = posh[active ? t : (t1 - 1)];
    const float a2 = pi.w * pi.w, ainv = 1.0f / pi.w;
    const float band = theta2 * 8.0e-6f;
    WalkAcc w = {0.f, 0.f, 0.f, 0.f, 0, 0};
    const unsigned m0 = __ballot_sync(FULL, active);
    if (m0 == 0) return;
    int4* st = stack[wid];
    int sp = 0;
    {   // root
        const float4 A = __ldg(&packed[0]), B = __ldg(&packed[1]);
        float dx, dy, dz, r_sq;
        const bool acc = active && mac_accept(pi, A, B.x, theta2, band, dx, dy, dz, r_sq);
        if (acc) m2p(w, A, dx, dy, dz, r_sq);
        const unsigned rej = __ballot_sync(FULL, active && !acc);
        const int ia = __float_as_int(B.y), ib = __float_as_int(B.z);
        if (rej != 0) {
            if (ib < 0) {
                const bool open = (rej >> lane) & 1u;
                const bool soft = __any_sync(FULL, open && r_sq < 2.0f * (a2 + B.x));
                p2p_bucket(w, pi, a2, ainv, posm, ia, -ib, open, soft);
            } else {
                if (lane == 0) st[0] = make_int4(ia, ib, (int)rej, 0);
                sp = 1;
            }
        }
        __syncwarp();
    }
    while (sp > 0) {
        const int4 e = st[--sp];
        __syncwarp();
        const float4 A0 = __ldg(&packed[2 * (size_t)e.x]), B0 = __ldg(&packed[2 * (size_t)e.x + 1]);
        const float4 A1 = __ldg(&packed[2 * (size_t)e.y]), B1 = __ldg(&packed[2 * (size_t)e.y + 1]);
        const bool mine = ((unsigned)e.z >> lane) & 1u;
        float dx0, dy0, dz0, r0, dx1, dy1, dz1, r1;
        const bool acc0 = mac_accept(pi, A0, B0.x, theta2, band, dx0, dy0, dz0, r0) && mine;
        const bool acc1 = mac_accept(pi, A1, B1.x, theta2, band, dx1, dy1, dz1, r1) && mine;
        if (acc0) m2p(w, A0, dx0, dy0, dz0, r0);
        if (acc1) m2p(w, A1, dx1, dy1, dz1, r1);
        const unsigned rej0 = __ballot_sync(FULL, mine && !acc0);
        const unsigned rej1 = __ballot_sync(FULL, mine && !acc1);
        if (rej0 != 0) {
            const int ia = __float_as_int(B0.y), ib = __float_as_int(B0.z);
            if (ib < 0) {
                const bool open = (rej0 >> lane) & 1u;
                const bool soft = __any_sync(FULL, open && r0 < 2.0f * (a2 + B0.x));
                p2p_bucket(w, pi, a2, ainv, posm, ia, -ib, open, soft);
            } else {
                if (sp >= SPH_TREE_STACK) { if (lane == 0) atomicExch(&err[ERR_TREE_STACK], 1); break; }
                if (lane == 0) st[sp] = make_int4(ia, ib, (int)rej0, 0);
                sp++;
            }
        }
        if (rej1 != 0) {
            const int ia = __float_as_int(B1.y), ib = __float_as_int(B1.z);
            if (ib < 0) {
                const bool open = (rej1 >> lane) & 1u;
                const bool soft = __any_sync(FULL, open && r1 < 2.0f * (a2 + B1.x));
                p2p_bucket(w, pi, a2, ainv, posm, ia, -ib, open, soft);
            } else {
                if (sp >= SPH_TREE_STACK) { if (lane == 0) atomicExch(&err[ERR_TREE_STACK], 1); break; }
                if (lane == 0) st[sp] = make_int4(ia, ib, (int)rej1, 0);
                sp++;
            }
        }
        __syncwarp();
    }
    if (active) {
        grav[t] = make_float4(G * w.gx, G * w.gy, G * w.gz, G * w.gp);
        npart[t] = w.np;
        napprox[t] = w.na;
    }
}

}  // namespace

// LBVH topology + moments/boxes/packed nodes on `stream` (the handle's stream, or the auxiliary stream when the build is
// overlapped with the neighbor pass).
int sph_launch_tree_build(sphb200_ctx* c, float dt, cudaStream_t stream) {
    int n = (int)c->n;
    if (n <= 0) return SPH_OK;
    SPH_CK(c, cudaMemsetAsync(c->flag, 0, (size_t)n * sizeof(int32_t), stream));
    k_lbvh_topology<<<sph_div_up(n, 256), 256, 0, stream>>>(c->keys[1], n, c->child, c->range, c->parent);
    SPH_LAUNCH_CHECK(c);
    k_lbvh_nodes<<<sph_div_up(2 * (int64_t)n - 1, 256), 256, 0, stream>>>(c->posh[c->cur], c->velm[c->cur], n, c->child, c->range,
                                                                         c->parent, c->p.leaf_max, c->p.aabb_mode, dt, c->flag, c->mom,
                                                                         c->nlo, c->nhi, c->packed);
    SPH_LAUNCH_CHECK(c);
    c->tree_valid = true;
    return SPH_OK;
}

int sph_launch_tree_walk(sphb200_ctx* c) {
    int n = (int)c->n;
    int t0 = (int)c->t0;
    int t1 = (c->t1 < 0 || c->t1 > c->n) ? n : (int)c->t1;
    int nt = t1 - t0;
    if (n <= 0 || nt <= 0) return SPH_OK;
    float theta2 = c->p.theta * c->p.theta;  // fp32 product, as k_Theta*k_Theta (GravityFieldSystem.cs:246)
    k_tree_walk<<<sph_div_up(nt, TW_WARPS * 32), TW_WARPS * 32, 0, c->stream>>>(c->posh[c->cur], c->posm, c->packed, t0, t1, theta2,
                                                                               c->p.G, c->grav, c->npart, c->napprox, c->err_d);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}
