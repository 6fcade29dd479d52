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
// Walk mapping: one warp = 32 consecutive (spatially coherent) sorted targets sharing ONE work stack of
// (node, mask of lanes that opened every ancestor).  Each step pops up to 32 entries and turns the work sideways:
//   1. lanes = NODES: every lane tests its node against all 32 targets (positions broadcast from shared memory) with the
//      exact reference MAC and gets a 32-bit accept mask -- 32x32 exact decisions in ~350 instructions;
//   2. two 32x32 bit transposes (butterfly shuffles) hand every lane = TARGET the set of batch nodes it accepts and the
//      set of leaf buckets it must open; internal nodes that some lane rejected push both children with that lane mask;
//   3. the opened buckets' bodies are flattened into a shared-memory list of body pairs and summed warp-wide (uniform loads; the
//      (body x lane) mask matrix is transposed 32 bodies at a time, so a lane holds one word of "my bodies" and a chunk is a
//      straight-line sequence of 16 packed pair evaluations entered by a switch; capped Newtonian law for every listed body, the
//      few inside a = h_i get the softened remainder afterwards): measured faster than lane-private P2P, whose per-batch imbalance
//      left 25 % of the lanes busy; lanes = TARGETS for M2P: every lane sums only the nodes it accepts itself (the shared-walk
//      union was 4.5x the per-lane M2P work), as many per batch as the average lane needs -- leftovers wait, with the batch's
//      nodes, in the other half of a double buffer.  Both loops use packed FP32 (two bodies / nodes per instruction).
// Every lane still sees exactly the accepted nodes / opened buckets of its private depth-first walk (per-particle MAC,
// numParticles / numApprox identical to the oracle); only the order in which a lane adds its contributions differs.
// The MAC "bmax_sq / r_sq < theta^2" is monotone in r_sq, so the build stores per node the exact threshold T with
// accept <=> r_sq > T (found by stepping ulps around bmax_sq/theta^2 with the IEEE division): no division in the walk.
#include "ctx.cuh"
#include <math.h>

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ int delta_fn(const uint32_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    return a == b ? 32 + __clz((uint32_t)i ^ (uint32_t)j) : __clz(a ^ b);
}

// Karras (2012) binary radix tree over (key, slot) -- integer-only, identical to orc_lbvh_topology.
// This launch builds the nodes of slots [i0, i1) (a group rank: its own Morton range; keys[] is the global key array).  An
// internal node whose particle range leaves [i0, i1) is a TOP node: its topology is also appended to `top` so that every rank
// can finish it once the ranks have exchanged their frontier moments (k_top_tree, kernels_group.cu).
__global__ void __launch_bounds__(256) k_lbvh_topology(const uint32_t* __restrict__ keys, int n, int i0, int i1, int2* __restrict__ child,
                                                       int2* __restrict__ range, int32_t* __restrict__ parent,
                                                       TopNode* __restrict__ top, int32_t* __restrict__ top_count, int32_t* __restrict__ err) {
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
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
    SPH_DBG_IDX(lc, 2 * (long long)n - 1); SPH_DBG_IDX(rc, 2 * (long long)n - 1);
    child[i] = make_int2(lc, rc);
    range[i] = make_int2(lo, hi);
    parent[lc] = i;
    parent[rc] = i;
    if (lo < i0 || hi >= i1) {
        const int k = atomicAdd(&top_count[0], 1);
        if (k < SPH_TOP_CAP) top[k] = TopNode{i, lc, rc, lo, hi, 0, 0, 0};
        else atomicExch(&err[ERR_TOP_TREE], 1);
    }
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

// Exact MAC threshold: accept <=> RN(b_sq / r_sq) < theta2 (AcceptApproximation :246).  The quotient is non-increasing in
// r_sq, so T = the largest float with RN(b_sq / T) >= theta2 gives accept <=> r_sq > T.  b_sq == 0: 0/r_sq = 0 accepts for
// every r_sq > 0 and 0/0 = NaN rejects => T = 0.
__device__ __forceinline__ float mac_threshold(float b_sq, float theta2) {
    if (!(b_sq > 0.0f)) return 0.0f;
    uint32_t u = __float_as_uint(__fdiv_rn(b_sq, theta2));
    while (u > 0u && !(__fdiv_rn(b_sq, __uint_as_float(u)) >= theta2)) --u;
    while (__fdiv_rn(b_sq, __uint_as_float(u + 1u)) >= theta2) ++u;
    return __uint_as_float(u);
}

// Packed walk node: [2k] = (cm.xyz, T) test record, [2k+1] = (M, a, b, Bmax^2) with (a,b) = (left,right) for nodes that
// are descended and (first, -count) for leaf buckets.  Bmax^2 with the op order of AcceptApproximation (:235-243).
__device__ __forceinline__ void pack_node(float4* __restrict__ packed, int k, float4 mo, const float lo[3], const float hi[3],
                                          int2 ch, int2 rg, int leaf_max, float theta2) {
    float bx = fmaxf(__fsub_rn(hi[0], mo.x), __fsub_rn(mo.x, lo[0]));
    float by = fmaxf(__fsub_rn(hi[1], mo.y), __fsub_rn(mo.y, lo[1]));
    float bz = fmaxf(__fsub_rn(hi[2], mo.z), __fsub_rn(mo.z, lo[2]));
    float b_sq = dot3_rn(bx, by, bz);
    int cnt = rg.y - rg.x + 1;
    int a = cnt <= leaf_max ? rg.x : ch.x;
    int b = cnt <= leaf_max ? -cnt : ch.y;
    packed[2 * (size_t)k] = make_float4(mo.x, mo.y, mo.z, mac_threshold(b_sq, theta2));
    packed[2 * (size_t)k + 1] = make_float4(mo.w, __int_as_float(a), __int_as_float(b), b_sq);
}

// One body of a bucket: GravitationalMoment.Accumulate (GravityFieldSystem.cs:398-411) and the union of the collider AABBs
// (quirk Q2) -- or of the positions, aabb_mode 1.
__device__ __forceinline__ void sph_bucket_add(float4& mo, float lo[3], float hi[3], const float4 p, const float4 v, int aabb_mode, float dt) {
    const float margin = 0.1f * 0.5f;  // CollisionTolerance * 0.5 (Broadphase.cs:200, CollisionWorld.cs:32)
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

// Moments + MAC boxes.  Buckets (maximal nodes with <= leaf_max bodies) are evaluated directly from their particle range
// (reference leaf rule; the nodes inside a bucket are unreachable and skipped); larger nodes are finished bottom-up by the second thread to arrive (fixed left-then-right order).
// nodes of this launch: the internal nodes g0 .. g1-1 and the leaves of slots g0 .. g1-1 (node ids n-1+slot); the particle
// of global slot s is resident at posh[s + off].  Single handle: g0 = 0, g1 = n, off = 0.  A group rank finishes only the nodes
// whose particle range lies inside [g0, g1); every finished node whose parent is unknown here (built by another rank) or
// straddles a rank boundary is appended to `front`.
__global__ void __launch_bounds__(256) k_lbvh_nodes(const float4* __restrict__ posh, const float4* __restrict__ velm, int n, int g0, int g1,
                                                    int off, const int2* __restrict__ child, const int2* __restrict__ range,
                                                    const int32_t* __restrict__ parent, int leaf_max, int aabb_mode, float dt,
                                                    float theta2, int32_t* __restrict__ flag, float4* mom, float4* nlo, float4* nhi,
                                                    float4* __restrict__ packed, FrontNode* __restrict__ front,
                                                    int32_t* __restrict__ top_count, int32_t* __restrict__ err) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, nown = g1 - g0;
    if (u >= 2 * nown) return;
    const int k = u < nown ? g0 + u : n - 1 + g0 + (u - nown);
    if (u < nown && k >= n - 1) return;   // n slots have n-1 internal nodes
    int2 rg = range[k];
    SPH_DBG_IDX(rg.x, n); SPH_DBG_IDX(rg.y, n);
    if (rg.y - rg.x + 1 > leaf_max) return;
    {   // a small node whose parent is small too lies strictly inside a bucket: the walk never reaches it
        const int p = parent[k];
        if (p >= 0) {
            const int2 prg = range[p];
            if (prg.y - prg.x + 1 <= leaf_max) return;
        }
    }
    if (rg.x < g0 || rg.y >= g1) return;  // bucket across a rank boundary: a top node (k_lbvh_topology listed it), finished by
                                          // k_top_tree from the ranks' boundary particles
    float4 mo = make_float4(0.f, 0.f, 0.f, 0.f);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int s = rg.x; s <= rg.y; s++) sph_bucket_add(mo, lo, hi, posh[s + off], velm[s + off], aabb_mode, dt);
    mom[k] = mo;
    nlo[k] = make_float4(lo[0], lo[1], lo[2], __int_as_float(rg.x));
    nhi[k] = make_float4(hi[0], hi[1], hi[2], __int_as_float(rg.y));
    pack_node(packed, k, mo, lo, hi, child[k], rg, leaf_max, theta2);
    int cur = k;
    float4 cmo = mo, clo = make_float4(lo[0], lo[1], lo[2], 0.f), chi = make_float4(hi[0], hi[1], hi[2], 0.f);
    while (true) {
        int p = parent[cur];
        SPH_DBG_IDX(p + 1, 2 * (long long)n);
        int2 prg = p >= 0 ? range[p] : make_int2(0, 0);
        if (p < 0 || prg.x < g0 || prg.y >= g1) {
            // the root -- or, on a group rank, a parent that is built elsewhere / straddles the rank boundary: `cur` is a
            // frontier node of the top tree
            if (cur != 0) {
                const int q = atomicAdd(&top_count[1], 1);
                if (q < SPH_TOP_CAP) front[q] = FrontNode{cur, 0, 0, 0, cmo, clo, chi};
                else atomicExch(&err[ERR_TOP_TREE], 1);
            }
            break;
        }
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
        pack_node(packed, p, acc, plo, phi, ch, prg, leaf_max, theta2);
        cur = p;
        cmo = acc; clo = make_float4(plo[0], plo[1], plo[2], 0.f); chi = make_float4(phi[0], phi[1], phi[2], 0.f);
    }
}

constexpr int TW_WARPS = 4;
constexpr int TW_STACK = 448;    // per-warp work stack entries (node, lane mask); 448 keeps the block at 30 KB of shared memory = 7 blocks per SM
constexpr int TW_HEADROOM = 72;  // > the deepest possible LBVH (30 key bits + 32 tie-break bits): once the stack is this close to
                                 // full the walk degrades to one node per step -- plain depth-first, which grows the stack by
                                 // at most one entry per tree level -- instead of failing (clustered / duplicate-key inputs)

struct WalkAcc {
    float gx, gy, gz, gp;
    int np, na;
};

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }


// Inside a = h_i the pair law is Dyer & Ip's (GravityFieldSystem.cs:340-347): g = (m/a^3)(8 - 9x + 2x^3), phi = -(m/a)(2.4 - 4x^2 +
// 3x^3 - 0.4x^5), x = r/a.  The packed loop has already summed the capped Newtonian value (m/a^3, -m/a) for the pair: this adds the rest.
__device__ __forceinline__ void p2p_soft(WalkAcc& w, float ex, float ey, float ez, float r2, float m, float ainv) {
    const float r = r2 > 0.f ? r2 * rsqrt_approx(r2) : 0.f;
    const float x = r * ainv, x2 = x * x, x3 = x2 * x;
    const float ma = m * ainv;
    const float g = ma * ainv * ainv * (7.0f - 9.0f * x + 2.0f * x3);
    w.gx = fmaf(ex, g, w.gx); w.gy = fmaf(ey, g, w.gy); w.gz = fmaf(ez, g, w.gz);
    w.gp -= ma * (1.4f - 4.0f * x2 + 3.0f * x3 - 0.4f * x2 * x3);
}

// One body of the packed P2P sequence: m <- (W & BIT) ? m : 0, and BIT is recorded in `soft` when the body is the lane's and lies
// inside the softening radius.  PTX so that the bit test stays one LOP3 with a predicate result (the compiler's own lowering of
// (W & BIT) != 0 is a shift, a mask and a compare per body).
template <unsigned BIT>
__device__ __forceinline__ void tw_mask_body(unsigned W, float& m, float r2, float a2, unsigned& soft) {
    asm("{\n\t.reg .pred p, q;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %2, %5;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "selp.f32 %0, %0, 0f00000000, p;\n\t"
        "setp.lt.and.f32 q, %3, %4, p;\n\t"
        "@q or.b32 %1, %1, %5;\n\t}"
        : "+f"(m), "+r"(soft) : "r"(W), "f"(r2), "f"(a2), "n"(BIT));
}

// One M2P iteration of a lane: up to two of its pending nodes, the previous batch's (mask mo, nodes bo) before this batch's (mc, bc).
__device__ __forceinline__ void tw_m2p_step(unsigned& mo, unsigned& mc, const float4* __restrict__ bo, const float4* __restrict__ bc,
                                            u64 pix, u64 piy, u64 piz, u64& gx2, u64& gy2, u64& gz2, u64& gp2) {
    const bool f0 = mo != 0u;
    unsigned m = f0 ? mo : mc;
    if (m == 0u) return;
    const float4* p0 = (f0 ? bo : bc) + (__ffs(m) - 1);
    m &= m - 1u;
    if (f0) mo = m; else mc = m;
    const bool f1 = mo != 0u;
    m = f1 ? mo : mc;
    const bool two = m != 0u;
    const float4* p1 = two ? (f1 ? bo : bc) + (__ffs(m) - 1) : p0;
    m &= m - 1u;                       // 0 & 0xffffffff = 0 when there was no second node
    if (f1) mo = m; else mc = m;
    const float4 A0 = *p0, A1 = *p1;
    const u64 dx = sub2(pix, pk2(A0.x, A1.x)), dy = sub2(piy, pk2(A0.y, A1.y)), dz = sub2(piz, pk2(A0.z, A1.z));
    const u64 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    float ra, rb;
    upk2(r2, ra, rb);
    const u64 rinv = pk2(rsqrt_approx(ra), rsqrt_approx(rb));
    const u64 mr = mul2(pk2(A0.w, two ? A1.w : 0.f), rinv);
    const u64 g = mul2(mr, mul2(rinv, rinv));
    gx2 = fma2(dx, g, gx2); gy2 = fma2(dy, g, gy2); gz2 = fma2(dz, g, gz2);
    gp2 = sub2(gp2, mr);
}

// 32x32 bit-matrix transpose across the warp: in: lane i holds row i, out: lane j holds column j
__device__ __forceinline__ unsigned transpose32(unsigned x, int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned m = o == 16 ? 0x0000FFFFu : o == 8 ? 0x00FF00FFu : o == 4 ? 0x0F0F0F0Fu : o == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(FULL, x, o);
        x = (lane & o) ? ((x & ~m) | ((y >> o) & m)) : ((x & m) | ((y << o) & ~m));
    }
    return x;
}

// Slots are GLOBAL sorted slots: posm / packed span all n particles (a group rank holds the all-gathered arrays), the targets
// are slots [t0, t1) and the resident arrays (posh, grav, ...) hold slot t at t + off.
__global__ void __launch_bounds__(TW_WARPS * 32) k_tree_walk(const float4* __restrict__ posh, const float4* __restrict__ posm,
                                                             const float4* __restrict__ packed, int n, int t0, int t1, int off, float G,
                                                             float negzero, float4* __restrict__ grav, int32_t* __restrict__ npart,
                                                             int32_t* __restrict__ napprox, int32_t* __restrict__ err) {
    __shared__ int2 stack[TW_WARPS][TW_STACK];
    __shared__ ulonglong2 tgxy[TW_WARPS][16];  // target positions as pairs: (x0,x1), (y0,y1)   (lanes = nodes phase)
    __shared__ u64 tgzz[TW_WARPS][16];         //                           (z0,z1)
    __shared__ float4 bcm[TW_WARPS][2][32];  // batch nodes: (cm, M) of this batch and of the previous one (lanes = targets phase)
    __shared__ __align__(16) float sbodyf[TW_WARPS][128 * 4];   // flattened bodies of the shared buckets, as pairs: (x0,x1,y0,y1),(z0,z1,m0,m1)
    __shared__ __align__(8) unsigned sbodym[TW_WARPS][128 + 2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // Warps are aligned to absolute multiples of 32 sorted slots and every slot < n walks, whether or not it lies in this
    // rank's target range [t0,t1): the order in which a lane adds its contributions depends on its 31 companions, so a
    // sharded run stays bit-identical to the single-GPU run only if the groups are the same.  Companions outside the range
    // take their position from the global source array (their softening length is irrelevant: the results are dropped).
    const int t = (t0 & ~31) + (blockIdx.x * TW_WARPS + wid) * 32 + lane;
    if (t - lane >= t1) return;   // whole group beyond the range
    const bool active = t < n;
    const bool mine = t >= t0 && t < t1;
    float4 pi = __ldg(&posm[active ? t : (n - 1)]);
    pi.w = mine ? posh[t + off].w : 1.0f;
    const float a2 = pi.w * pi.w, ainv = 1.0f / pi.w;
    WalkAcc w = {0.f, 0.f, 0.f, 0.f, 0, 0};
    u64 gx2 = pk2(0.f, 0.f), gy2 = gx2, gz2 = gx2, gp2 = gx2;   // packed partial sums (P2P and M2P loops)
    const u64 pix = pk2(pi.x, pi.x), piy = pk2(pi.y, pi.y), piz = pk2(pi.z, pi.z);
    const unsigned m0 = __ballot_sync(FULL, active);
    if (m0 == 0) return;
    int2* st = stack[wid];
    ulonglong2* tgp = tgxy[wid];
    u64* tgz = tgzz[wid];
    float4* bcb = &bcm[wid][0][0];
    unsigned mold = 0u;   // accepted nodes of the previous batch this lane has not summed yet (M2P below)
    int par = 0;
    float* sbf = sbodyf[wid];
    unsigned* sbm = sbodym[wid];
    {
        float* fx = reinterpret_cast<float*>(tgp);
        float* fz = reinterpret_cast<float*>(tgz);
        fx[(lane >> 1) * 4 + (lane & 1)] = pi.x;
        fx[(lane >> 1) * 4 + 2 + (lane & 1)] = pi.y;
        fz[lane] = pi.z;
    }
    if (lane == 0) st[0] = make_int2(0, (int)m0);
    int sp = 1;
    __syncwarp();
    while (sp > 0) {
        // ---- 1. lanes = nodes
        // a batch of nb nodes grows the stack by at most nb; wide batches keep TW_HEADROOM entries free
        const int nb = max(min(min(sp, 32), TW_STACK - TW_HEADROOM - sp), 1);
        if (sp >= TW_STACK) { if (lane == 0) atomicExch(&err[ERR_TREE_STACK], 1); break; }   // unreachable for trees of depth < TW_HEADROOM
        const bool have = lane < nb;
        int2 e = make_int2(0, 0);
        if (have) { SPH_DBG_IDX(sp - 1 - lane, TW_STACK); e = st[sp - 1 - lane]; SPH_DBG_IDX(e.x, 2 * (long long)n - 1); }
        sp -= nb;
        const float4 N = __ldg(&packed[2 * (size_t)e.x]);
        const float4 X = __ldg(&packed[2 * (size_t)e.x + 1]);
        // exact MAC, 2 targets per packed subtract / multiply (FADD2 / FMUL2 are IEEE round-to-nearest per half): the
        // reference's (dx*dx + dy*dy) + dz*dz bit for bit
        // Squares as fma(d, d, -0) with a -0 the compiler cannot see (a kernel argument): ptxas contracts mul.rn.f32x2 +
        // add.rn.f32x2 into one FFMA2 (a single rounding) even with explicit .rn, but it cannot contract two fma results, so
        // (dx*dx + dy*dy) + dz*dz keeps the reference's three roundings while every operation stays packed.
        unsigned amask = 0u;
        const u64 nx = pk2(N.x, N.x), ny = pk2(N.y, N.y), nz = pk2(N.z, N.z), tw = pk2(N.w, N.w), z2 = pk2(negzero, negzero);
#pragma unroll
        for (int tp = 15; tp >= 0; tp--) {         // descending: every result is shifted in at bit 0
            const ulonglong2 P = tgp[tp];          // (x0,x1), (y0,y1)
            const u64 Z = tgz[tp];                 // (z0,z1)
            const u64 dx = sub2(P.x, nx), dy = sub2(P.y, ny), dz = sub2(Z, nz);
            const u64 r2 = add2(add2(fma2(dx, dx, z2), fma2(dy, dy, z2)), fma2(dz, dz, z2));
            // AcceptApproximation, exact: r_sq > T <=> T - r_sq < 0 (a difference of distinct floats never rounds to zero)
            float sa, sb;
            upk2(sub2(tw, r2), sa, sb);
            amask = __funnelshift_l(__float_as_uint(sb), amask, 1);
            amask = __funnelshift_l(__float_as_uint(sa), amask, 1);
        }
        const unsigned mask = (unsigned)e.y;
        const unsigned acc = amask & mask, rej = mask & ~amask;
        const int ia = __float_as_int(X.y), ib = __float_as_int(X.z);
        const bool bucket = ib < 0;
        __syncwarp();
        float4* bc = bcb + 32 * par;            // this batch's nodes; the other half holds the previous batch's
        const float4* bo = bcb + 32 * (par ^ 1);
        bc[lane] = make_float4(N.x, N.y, N.z, X.x);
        // internal nodes some lane rejected: both children inherit that lane mask
        const bool open = have && !bucket && rej != 0u;
        const unsigned ob = __ballot_sync(FULL, open);
        if (open) {
            const int pos = sp + 2 * __popc(ob & ((1u << lane) - 1u));
            SPH_DBG_IDX(pos + 1, TW_STACK);
            SPH_DBG_IDX(ia, 2 * (long long)n - 1); SPH_DBG_IDX(ib, 2 * (long long)n - 1);
            st[pos] = make_int2(ia, (int)rej);
            st[pos + 1] = make_int2(ib, (int)rej);
        }
        sp += 2 * __popc(ob);
        // ---- 2. sideways: which batch nodes does target `lane` accept
        unsigned nmask = transpose32(acc, lane);
        w.na += __popc(nmask);
        __syncwarp();
        // ---- 3a. P2P, warp-wide (GravityContributionParticle :332-356, a = h_i; includes the target itself, Q3): the bodies of
        // the opened buckets are flattened (4 per bucket per round; one round when leaf_max <= 4) into a shared-memory list of
        // body PAIRS, SoA inside the pair like the all-pairs tiles, and summed with packed FP32: one instruction, two bodies
        {
            int brem = (have && bucket && rej != 0u) ? -ib : 0, bfirst = ia;
            while (__any_sync(FULL, brem > 0)) {
                const int c4 = min(brem, 4);
                int incl = c4;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(FULL, incl, 31);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (u < c4) {
                        const int pos = incl - c4 + u;
                        SPH_DBG_IDX(pos, 128);
                        SPH_DBG_IDX(bfirst + u, n);
                        const float4 pj = __ldg(&posm[bfirst + u]);
                        float* rec = sbf + (pos >> 1) * 8 + (pos & 1);
                        rec[0] = pj.x; rec[2] = pj.y; rec[4] = pj.z; rec[6] = pj.w;
                        sbm[pos] = rej;
                    }
                if (lane == 0 && (total & 1)) {   // odd count: the last pair's second body is a far-away zero-mass dummy
                    float* rec = sbf + (total >> 1) * 8 + 1;
                    rec[0] = 1.0e18f; rec[2] = 1.0e18f; rec[4] = 1.0e18f; rec[6] = 0.f;
                    sbm[total] = 0u;
                }
                bfirst += c4; brem -= c4;
                __syncwarp();
                const ulonglong2* recs = reinterpret_cast<const ulonglong2*>(sbf);
                // The (body x lane) mask matrix is turned sideways 32 bodies at a time: lane = target then holds one word whose
                // bit k says "body k of the chunk is mine" -- no mask load or variable shift per body, and numParticles is one
                // popc per chunk.  Every listed body is summed with the all-pairs kernel's capped Newtonian law (r capped at
                // a = h_i, kernels_gravity.cu); the few bodies inside a get "Dyer & Ip minus capped" in the branch below.
                for (int c0 = 0; c0 < total; c0 += 32) {
                    unsigned W = transpose32(c0 + lane < total ? sbm[c0 + lane] : 0u, lane);
                    w.np += __popc(W);
                    // The chunk's pairs run through one straight-line sequence of 16 pair evaluations entered at 16 - npair
                    // (fall-through switch: no exit test and no shift per pair); pair q of the sequence is the chunk's pair
                    // q - (16 - npair), so the mask word and the record pointer are aligned to the END of the sequence.
                    const int npair = (min(total - c0, 32) + 1) >> 1, skip = 16 - npair;
                    const ulonglong2* rp = recs + c0 - 2 * skip;
                    W <<= 2 * skip;
                    unsigned soft = 0u;   // bodies inside the softening radius (rare: finished after the sequence)
#define TW_PAIR(q)                                                                                                              \
                    {                                                                                                           \
                        const ulonglong2 A = rp[2 * (q)], B = rp[2 * (q) + 1]; /* (x0,x1),(y0,y1) | (z0,z1),(m0,m1) */          \
                        const u64 ex = sub2(pix, A.x), ey = sub2(piy, A.y), ez = sub2(piz, B.x);                                \
                        const u64 r2 = fma2(ez, ez, fma2(ey, ey, mul2(ex, ex)));                                                \
                        float ra, rb, ma, mb;                                                                                   \
                        upk2(r2, ra, rb);                                                                                       \
                        upk2(B.y, ma, mb);                                                                                      \
                        tw_mask_body<(1u << (2 * (q)))>(W, ma, ra, a2, soft);                                                   \
                        tw_mask_body<(2u << (2 * (q)))>(W, mb, rb, a2, soft);                                                   \
                        const u64 rinv = pk2(rsqrt_approx(fmaxf(ra, a2)), rsqrt_approx(fmaxf(rb, a2)));                         \
                        const u64 mr = mul2(pk2(ma, mb), rinv);                                                                 \
                        const u64 g = mul2(mr, mul2(rinv, rinv));                                                               \
                        gx2 = fma2(ex, g, gx2); gy2 = fma2(ey, g, gy2); gz2 = fma2(ez, g, gz2);                                 \
                        gp2 = sub2(gp2, mr);                                                                                    \
                    }
                    switch (skip) {
                        case 0: TW_PAIR(0)
                        case 1: TW_PAIR(1)
                        case 2: TW_PAIR(2)
                        case 3: TW_PAIR(3)
                        case 4: TW_PAIR(4)
                        case 5: TW_PAIR(5)
                        case 6: TW_PAIR(6)
                        case 7: TW_PAIR(7)
                        case 8: TW_PAIR(8)
                        case 9: TW_PAIR(9)
                        case 10: TW_PAIR(10)
                        case 11: TW_PAIR(11)
                        case 12: TW_PAIR(12)
                        case 13: TW_PAIR(13)
                        case 14: TW_PAIR(14)
                        default: TW_PAIR(15)
                    }
#undef TW_PAIR
                    soft >>= 2 * skip;
                    while (soft) {
                        const int b = __ffs(soft) - 1;
                        soft &= soft - 1u;
                        const float* rec = sbf + ((c0 + b) >> 1) * 8 + (b & 1);
                        const float ex = pi.x - rec[0], ey = pi.y - rec[2], ez = pi.z - rec[4];
                        p2p_soft(w, ex, ey, ez, fmaf(ez, ez, fmaf(ey, ey, ex * ex)), rec[6], ainv);
                    }
                }
                __syncwarp();
            }
        }
        // ---- 3b. M2P, lane-private (GravitationalMoment.GravityContribution, :428-442): every lane sums the nodes it accepts, two
        // per iteration with packed FP32 (an odd last one is paired with a zero-mass copy of itself; r_sq > T >= 0).  The lanes'
        // sets differ in size (a batch leaves 49 % of the lane-iterations idle if every lane must finish its own set), so a batch
        // runs only as many iterations as the AVERAGE lane needs -- and as the previous batch's leftovers need, which go first --
        // and what a lane has left of this batch waits, with the batch's nodes in the other half of the buffer, for the next one.
        {
            const int po = __popc(mold), pn = __popc(nmask);
            const int T = max((__reduce_max_sync(FULL, po) + 1) >> 1, (__reduce_add_sync(FULL, po + pn) + 63) >> 6);
            for (int it = T; it > 0; it--) tw_m2p_step(mold, nmask, bo, bc, pix, piy, piz, gx2, gy2, gz2, gp2);
            mold = nmask;      // T covered every lane's old set
            par ^= 1;
        }
        __syncwarp();
    }
    {   // what is left of the last batch
        const float4* bo = bcb + 32 * (par ^ 1);
        unsigned none = 0u;
        for (int it = (__reduce_max_sync(FULL, __popc(mold)) + 1) >> 1; it > 0; it--) tw_m2p_step(mold, none, bo, bo, pix, piy, piz, gx2, gy2, gz2, gp2);
    }
    if (mine) {
        float a, b;
        upk2(gx2, a, b); w.gx += a + b;
        upk2(gy2, a, b); w.gy += a + b;
        upk2(gz2, a, b); w.gz += a + b;
        upk2(gp2, a, b); w.gp += a + b;
        grav[t + off] = make_float4(G * w.gx, G * w.gy, G * w.gz, G * w.gp);
        npart[t + off] = w.np;
        napprox[t + off] = w.na;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Top tree of the distributed build (group ranks only).  Inputs, all-gathered from every rank: the topology of the nodes
// that straddle a rank boundary (`top`), the moments/boxes of the finished nodes hanging below them (`front`) and the
// first / last TOP_LEAF particles of every rank (`bnd`, for buckets that straddle a boundary).  One block finishes the
// straddling nodes bottom-up, sweep by sweep, with the reference's arithmetic in the reference's order -- every rank
// computes the same bits.  Writes mom / nlo / nhi / packed of the top nodes.
// ------------------------------------------------------------------------------------------------------------
constexpr int TOP_MAX = 4096;    // straddling nodes of the whole group (<= (world-1) * tree depth)

struct TopArgs {
    const TopNode* top; const FrontNode* front; const int32_t* counts;   // [world][SPH_TOP_CAP], [world][SPH_TOP_CAP], [world][4]
    const float4* bnd;            // [world][2][SPH_TOP_LEAF][2]: (posh, velm) of the first / last particles (last: right-aligned)
    const int64_t* g0;            // [world+1] global slot ranges of the ranks (device)
    int world, leaf_max, aabb_mode; float dt, theta2;
    float4 *mom, *nlo, *nhi, *packed; int32_t* err;
};

__global__ void __launch_bounds__(1024) k_top_tree(TopArgs A) {
    __shared__ int tid_[TOP_MAX];              // node id of top node t
    __shared__ short trank[TOP_MAX];           // where its record lives
    __shared__ short tslot[TOP_MAX];
    __shared__ unsigned char done[TOP_MAX];
    __shared__ int nbase[SPH_MAX_RANKS + 1];
    const int tx = threadIdx.x;
    if (tx == 0) {
        int acc = 0;
        for (int r = 0; r < A.world; r++) { nbase[r] = acc; acc += min(A.counts[4 * r], SPH_TOP_CAP); }
        nbase[A.world] = acc;
    }
    __syncthreads();
    const int T = min(nbase[A.world], TOP_MAX);
    if (nbase[A.world] > TOP_MAX && tx == 0) atomicExch(&A.err[ERR_TOP_TREE], 1);
    for (int t = tx; t < T; t += blockDim.x) {
        int r = 0;
        while (nbase[r + 1] <= t) r++;
        trank[t] = (short)r; tslot[t] = (short)(t - nbase[r]);
        tid_[t] = A.top[r * SPH_TOP_CAP + (t - nbase[r])].id;
        done[t] = 0;
    }
    // frontier moments of every rank into the node arrays
    for (int r = 0; r < A.world; r++) {
        const int nf = min(A.counts[4 * r + 1], SPH_TOP_CAP);
        for (int k = tx; k < nf; k += blockDim.x) {
            const FrontNode f = A.front[r * SPH_TOP_CAP + k];
            A.mom[f.id] = f.mom; A.nlo[f.id] = f.lo; A.nhi[f.id] = f.hi;
        }
    }
    __syncthreads();
    // children of my nodes: index in the top list, or -1 (a frontier node: ready)
    constexpr int PER = TOP_MAX / 1024;
    int ca[PER], cb[PER];
#pragma unroll
    for (int u = 0; u < PER; u++) {
        const int t = tx + u * 1024;
        ca[u] = cb[u] = -1;
        if (t < T) {
            const TopNode nd = A.top[trank[t] * SPH_TOP_CAP + tslot[t]];
            if (nd.hi - nd.lo + 1 > A.leaf_max)
                for (int q = 0; q < T; q++) {
                    const int id = tid_[q];
                    if (id == nd.a) ca[u] = q;
                    if (id == nd.b) cb[u] = q;
                }
        }
    }
    for (int sweep = 0; sweep < 4 * TOP_MAX; sweep++) {
        int progress = 0;
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int t = tx + u * 1024;
            if (t >= T || done[t]) continue;
            const TopNode nd = A.top[trank[t] * SPH_TOP_CAP + tslot[t]];
            const int2 rg = make_int2(nd.lo, nd.hi);
            float4 mo = make_float4(0.f, 0.f, 0.f, 0.f);
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            if (nd.hi - nd.lo + 1 <= A.leaf_max) {
                // bucket across a boundary: its bodies are among the first / last particles of their ranks
                int r = 0;
                for (int s = nd.lo; s <= nd.hi; s++) {
                    while (A.g0[r + 1] <= s) r++;
                    const int i = s - (int)A.g0[r];
                    const int j = s - ((int)A.g0[r + 1] - SPH_TOP_LEAF);
                    const float4* rec = A.bnd + ((size_t)(r * 2 + (i < SPH_TOP_LEAF ? 0 : 1)) * SPH_TOP_LEAF + (i < SPH_TOP_LEAF ? i : j)) * 2;
                    sph_bucket_add(mo, lo, hi, rec[0], rec[1], A.aabb_mode, A.dt);
                }
            } else {
                if ((ca[u] >= 0 && done[ca[u]] != 1) || (cb[u] >= 0 && done[cb[u]] != 1)) continue;
                const float4 ml = __ldcg(&A.mom[nd.a]), mr = __ldcg(&A.mom[nd.b]);
                const float4 ll = __ldcg(&A.nlo[nd.a]), lr = __ldcg(&A.nlo[nd.b]);
                const float4 hl = __ldcg(&A.nhi[nd.a]), hr = __ldcg(&A.nhi[nd.b]);
                moment_accumulate(mo, ml.x, ml.y, ml.z, ml.w);  // GravityFieldSystem.cs:513-521, children in Data order
                moment_accumulate(mo, mr.x, mr.y, mr.z, mr.w);
                lo[0] = fminf(ll.x, lr.x); lo[1] = fminf(ll.y, lr.y); lo[2] = fminf(ll.z, lr.z);
                hi[0] = fmaxf(hl.x, hr.x); hi[1] = fmaxf(hl.y, hr.y); hi[2] = fmaxf(hl.z, hr.z);
            }
            A.mom[nd.id] = mo;
            A.nlo[nd.id] = make_float4(lo[0], lo[1], lo[2], __int_as_float(rg.x));
            A.nhi[nd.id] = make_float4(hi[0], hi[1], hi[2], __int_as_float(rg.y));
            pack_node(A.packed, nd.id, mo, lo, hi, make_int2(nd.a, nd.b), rg, A.leaf_max, A.theta2);
            __threadfence_block();
            done[t] = 2;      // visible as "done" only after the barrier below (value 2 -> 1)
            progress = 1;
        }
        const int any = __syncthreads_or(progress);
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int t = tx + u * 1024;
            if (t < T && done[t] == 2) done[t] = 1;
        }
        __syncthreads();
        if (!any) break;
    }
    // every top node must have been finished (otherwise a child record was missing: list overflow)
    int left = 0;
#pragma unroll
    for (int u = 0; u < PER; u++) {
        const int t = tx + u * 1024;
        if (t < T && !done[t]) left = 1;
    }
    if (left) atomicExch(&A.err[ERR_TOP_TREE], 1);
}

}  // namespace

// LBVH topology + moments/boxes/packed nodes on `stream` (the handle's stream, or the auxiliary stream when the build is
// overlapped with the neighbor pass).
// The tree spans c->tree_n global slots (keys c->tkeys); this context builds the nodes of slots [tree_g0, tree_g1).
int sph_launch_tree_build(sphb200_ctx* c, float dt, cudaStream_t stream) {
    const int n = (int)c->tree_n, g0 = (int)c->tree_g0, g1 = (int)c->tree_g1, own = g1 - g0;
    if (n <= 0 || own <= 0) return SPH_OK;
    const bool part = own < n;   // group rank: unknown parents must read as -1 (root / built elsewhere)
    TopNode* top = c->top_nodes ? c->top_nodes + (size_t)c->top_rank * SPH_TOP_CAP : nullptr;
    FrontNode* front = c->front_nodes ? c->front_nodes + (size_t)c->top_rank * SPH_TOP_CAP : nullptr;
    int32_t* tcount = c->top_counts ? c->top_counts + 4 * c->top_rank : nullptr;
    SPH_CK(c, cudaMemsetAsync(c->flag + g0, 0, (size_t)own * sizeof(int32_t), stream));
    if (part) {
        SPH_CK(c, cudaMemsetAsync(c->parent + g0, 0xff, (size_t)own * sizeof(int32_t), stream));
        SPH_CK(c, cudaMemsetAsync(c->parent + (n - 1 + g0), 0xff, (size_t)own * sizeof(int32_t), stream));
        SPH_CK(c, cudaMemsetAsync(tcount, 0, 4 * sizeof(int32_t), stream));
    }
    k_lbvh_topology<<<sph_div_up(own, 256), 256, 0, stream>>>(c->tkeys, n, g0, g1, c->child, c->range, c->parent, top, tcount, c->err_d);
    SPH_LAUNCH_CHECK(c);
    const float4* bposh = c->tree_posh ? c->tree_posh : c->posh[c->cur];
    const float4* bvelm = c->tree_posh ? c->tree_velm : c->velm[c->cur];
    const int boff = (int)(c->tree_posh ? c->tree_src_off : c->tree_off);
    k_lbvh_nodes<<<sph_div_up(2 * (int64_t)own, 256), 256, 0, stream>>>(bposh, bvelm, n, g0, g1, boff,
                                                                       c->child, c->range, c->parent, c->p.leaf_max, c->p.aabb_mode, dt,
                                                                       c->p.theta * c->p.theta /* fp32 product, as k_Theta*k_Theta (GravityFieldSystem.cs:246) */,
                                                                       c->flag, c->mom, c->nlo, c->nhi, c->packed, front, tcount, c->err_d);
    SPH_LAUNCH_CHECK(c);
    c->tree_valid = true;
    return SPH_OK;
}

// Finish the nodes that straddle rank boundaries (after the all-gather of top lists, frontier moments, boundary particles).
int sph_launch_top_tree(sphb200_ctx* c, int world, const float4* bnd, const int64_t* g0_d, float dt) {
    TopArgs A;
    A.top = c->top_nodes; A.front = c->front_nodes; A.counts = c->top_counts; A.bnd = bnd; A.g0 = g0_d;
    A.world = world; A.leaf_max = c->p.leaf_max; A.aabb_mode = c->p.aabb_mode; A.dt = dt; A.theta2 = c->p.theta * c->p.theta;
    A.mom = c->mom; A.nlo = c->nlo; A.nhi = c->nhi; A.packed = c->packed; A.err = c->err_d;
    k_top_tree<<<1, 1024, 0, c->stream>>>(A);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

int sph_launch_tree_walk(sphb200_ctx* c) {
    // resident target slots [t0, t1) -> global slots [t0 - off, t1 - off)
    int n = (int)c->tree_n, off = (int)c->tree_off;
    int t0 = (int)c->t0;
    int t1 = (c->t1 < 0 || c->t1 > c->n) ? (int)c->n : (int)c->t1;
    int nt = t1 - t0;
    if (n <= 0 || nt <= 0) return SPH_OK;
    t0 -= off; t1 -= off;
    k_tree_walk<<<sph_div_up(t1 - (t0 & ~31), TW_WARPS * 32), TW_WARPS * 32, 0, c->stream>>>(c->posh[c->cur], c->gsrc, c->packed, n, t0, t1, off,
                                                                               c->p.G, -0.0f, c->grav, c->npart, c->napprox, c->err_d);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}
