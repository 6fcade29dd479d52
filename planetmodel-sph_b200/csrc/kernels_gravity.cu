// kernels_gravity.cu -- all-pairs (direct-sum) self-gravity: shared-memory-tiled FP32 FMA kernel.
//
// Replaces: GravityFieldSystem.OnUpdateParticle + GravityContributionParticle
// (A/Systems/GravityFieldSystem.cs:249-303, 332-356): for every target i, sum over all j != i of
//   r <  a=h_i : d * (m/a^3)(8 - 9x + 2x^3),  phi = -(m/a)(2.4 - 4x^2 + 3x^3 - 0.4x^5),  x = r/a   (Dyer & Ip)
//   r >= a     : d * m/r^3,                   phi = -m/r
//
// B200 mapping: FP32 FMA-pipe bound (no tensor cores: not a contraction), so the design goal is the fewest issue
// slots per pair.  The inner loop is branch-free Newtonian:
//   * sources are Morton-sorted, so a 256-source tile is spatially compact; a tile whose box does not touch the box of
//     the block's targets grown by their softening radii is FAR (r >= h_i for every pair) and runs the bare loop;
//   * the few NEAR tiles (they also hold the self pair) run the same loop with r^2 capped from below at h_i^2 by one
//     FMNMX -- finite and smooth; the exact "softened minus capped" difference for pairs with r < h_i (all of them are
//     in i's SPH neighbor list, r < h_i < 2 max(h_i,h_j)) is added afterwards by k_gravity_near (kernels_sph.cu) and
//     the capped self term is removed in k_gravity_reduce;
//   * when every particle has the same mass (the reference spawner, ParticleAuthoring.cs:208) the mass multiply
//     leaves the loop (EQM variant) and is applied once in the reduce.
// Issue slots per pair (SASS): 7.4 far/equal-mass (packed FP32, below: per 2 sources x 6 targets 144 packed FMA-pipe instructions,
// 24 MUFU, 4 LDS, 5 loop instructions -- nothing but the formulation's 12 lane-ops per pair is left); rates in profiles/README.md.
// Summation: per-tile fp32 partials (256 sources) added into a running sum, then a fixed-order reduction over the
// source splits -- deterministic, and ~sqrt(256) times tighter than one 10^6-term fp32 chain.
#include "ctx.cuh"
#include <math.h>
#include <algorithm>

namespace {

constexpr int AP_THREADS = 256;
constexpr int AP_TPT = 6;     // targets per thread: 6 x unroll 2 (160 registers, ONE 8-warp CTA per SM, 12 independent pair chains per thread)
                              // measures 421 ms at 1M against 442 ms for 4 x 4 (128 registers, two CTAs); 3x4 440, 5x2 455, 8x2 464, 10x1 429,
                              // 6x1 489, 6x4 487, 7x2 534 (profiles/README.md, round 2): ILP per warp beats resident warps here
constexpr int AP_TILE = 256;  // sources per shared-memory tile (inner fp32 partial sums: 2 interleaved chains of 128)
constexpr int AP_UNROLL = 2;  // source pairs per unrolled inner-loop body
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Bounding box of every 256-source tile (2 float4 per tile: lo, hi).
__global__ void __launch_bounds__(AP_TILE) k_tile_boxes(const float4* __restrict__ src, int n, float4* __restrict__ tbox) {
    __shared__ float s[6][AP_TILE / 32];
    int i = min(blockIdx.x * AP_TILE + threadIdx.x, n - 1);
    float4 p = src[i];
    float v[6] = {p.x, p.y, p.z, p.x, p.y, p.z};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            v[k] = fminf(v[k], __shfl_xor_sync(FULL, v[k], o));
            v[3 + k] = fmaxf(v[3 + k], __shfl_xor_sync(FULL, v[3 + k], o));
        }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < 6; k++) s[k][w] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        float r[6];
        for (int k = 0; k < 6; k++) {
            r[k] = s[k][0];
            for (int j = 1; j < AP_TILE / 32; j++) r[k] = k < 3 ? fminf(r[k], s[k][j]) : fmaxf(r[k], s[k][j]);
        }
        tbox[2 * blockIdx.x] = make_float4(r[0], r[1], r[2], 0.f);
        tbox[2 * blockIdx.x + 1] = make_float4(r[3], r[4], r[5], 0.f);
    }
}

// Packed FP32 (Blackwell FFMA2 / FADD2 / FMUL2, PTX *.f32x2): one instruction evaluates the same target against TWO
// sources, halving the issue slots per pair (13.5 -> 7); the FMA pipe, not the scheduler, becomes the limit.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// One 256-source tile, stored as 128 source pairs: [2p] = (x0,x1,y0,y1), [2p+1] = (z0,z1,m0,m1).  The target
// coordinates are packed (xi,xi): ptxas turns them into the broadcast-scalar operand form (R.F32).
template <int TPT, bool CAP, bool EQM>
__device__ __forceinline__ void tile_loop(const ulonglong2* __restrict__ tp, const u64 (&xi)[TPT], const u64 (&yi)[TPT],
                                          const u64 (&zi)[TPT], const float (&a2)[TPT], float (&AX)[TPT], float (&AY)[TPT],
                                          float (&AZ)[TPT], float (&PH)[TPT]) {
    u64 ax[TPT], ay[TPT], az[TPT], ph[TPT];
#pragma unroll
    for (int k = 0; k < TPT; k++) ax[k] = ay[k] = az[k] = ph[k] = pk(0.f, 0.f);
#pragma unroll 1
    for (int j = 0; j < AP_TILE / 2; j += AP_UNROLL, tp += 2 * AP_UNROLL) {
#pragma unroll
        for (int u = 0; u < AP_UNROLL; u++) {
            const ulonglong2 A = tp[2 * u], B = tp[2 * u + 1];
#pragma unroll
            for (int k = 0; k < TPT; k++) {
                u64 dx = sub2(xi[k], A.x), dy = sub2(yi[k], A.y), dz = sub2(zi[k], B.x);
                u64 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                float r2a, r2b;
                upk(r2, r2a, r2b);
                if (CAP) { r2a = fmaxf(r2a, a2[k]); r2b = fmaxf(r2b, a2[k]); }
                u64 rinv = pk(rsqrt_approx(r2a), rsqrt_approx(r2b));
                u64 g;
                if (EQM) {
                    g = mul2(mul2(rinv, rinv), rinv);
                    ph[k] = add2(ph[k], rinv);
                } else {
                    u64 mr = mul2(B.y, rinv);
                    g = mul2(mr, mul2(rinv, rinv));
                    ph[k] = add2(ph[k], mr);
                }
                ax[k] = fma2(dx, g, ax[k]);
                ay[k] = fma2(dy, g, ay[k]);
                az[k] = fma2(dz, g, az[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < TPT; k++) {   // tile partial (2 x 128 sources) into the running sum
        float a, b;
        upk(ax[k], a, b); AX[k] += a + b;
        upk(ay[k], a, b); AY[k] += a + b;
        upk(az[k], a, b); AZ[k] += a + b;
        upk(ph[k], a, b); PH[k] += a + b;
    }
}

// Source tiles are double-buffered in shared memory (one barrier per tile; the next tile's global load is in flight
// during the math) and read back as warp-uniform LDS.128 broadcasts.
template <int TPT, bool EQM>
__global__ void __launch_bounds__(AP_THREADS) k_gravity_allpairs(const float4* __restrict__ src, int n_src, int src_per_split,
                                                                 const float4* __restrict__ tbox, const float4* __restrict__ posh,
                                                                 int t0, int t1, float4* __restrict__ part) {
    __shared__ __align__(16) float tile[2][AP_TILE * 4];
    __shared__ float4 tb_lo[2], tb_hi[2];
    __shared__ float ebox[6][AP_THREADS / 32];
    const int nt = t1 - t0;
    const int tid = threadIdx.x;
    const int tb = blockIdx.x * (AP_THREADS * TPT);
    u64 xi[TPT], yi[TPT], zi[TPT];
    float a2[TPT];
    float AX[TPT], AY[TPT], AZ[TPT], PH[TPT];
    float e[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * AP_THREADS + tid;
        float4 p = posh[t0 + min(t, nt - 1)];
        xi[k] = pk(p.x, p.x); yi[k] = pk(p.y, p.y); zi[k] = pk(p.z, p.z); a2[k] = p.w * p.w;
        AX[k] = AY[k] = AZ[k] = PH[k] = 0.f;
        // box of the block's targets grown by their softening radii (+0.1% against rounding)
        float a = p.w * 1.001f;
        e[0] = fminf(e[0], p.x - a); e[1] = fminf(e[1], p.y - a); e[2] = fminf(e[2], p.z - a);
        e[3] = fmaxf(e[3], p.x + a); e[4] = fmaxf(e[4], p.y + a); e[5] = fmaxf(e[5], p.z + a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            e[k] = fminf(e[k], __shfl_xor_sync(FULL, e[k], o));
            e[3 + k] = fmaxf(e[3 + k], __shfl_xor_sync(FULL, e[3 + k], o));
        }
    }
    if ((tid & 31) == 0)
        for (int k = 0; k < 6; k++) ebox[k][tid >> 5] = e[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float r = ebox[k][0];
#pragma unroll
        for (int j = 1; j < AP_THREADS / 32; j++) r = k < 3 ? fminf(r, ebox[k][j]) : fmaxf(r, ebox[k][j]);
        e[k] = r;
    }
    const int s0 = blockIdx.y * src_per_split;
    const int s1 = min(s0 + src_per_split, n_src);
    const int ntiles = (s1 - s0 + AP_TILE - 1) / AP_TILE;
    const int tile0 = s0 / AP_TILE;
    // padding source: zero mass (EQM: weight folded into position), far away: contributes exactly 0
    const float4 pad = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f);
    constexpr int LPT = AP_TILE / AP_THREADS;   // sources a thread stages per tile
    static_assert(AP_TILE % AP_THREADS == 0 && AP_TILE % (2 * AP_UNROLL) == 0, "tile shape");
    float4 nxt[LPT];
#pragma unroll
    for (int l = 0; l < LPT; l++) nxt[l] = (s0 + l * AP_THREADS + tid < s1) ? src[s0 + l * AP_THREADS + tid] : pad;
    float4 nbx = tid < 2 ? tbox[2 * tile0 + tid] : pad;
    for (int it = 0; it < ntiles; it++) {
        float* buf = tile[it & 1];
#pragma unroll
        for (int l = 0; l < LPT; l++) {
            const int sidx = l * AP_THREADS + tid;                 // source `sidx` of the tile
            const int so = (sidx >> 1) * 8 + (sidx & 1);          // its slot inside its pair record
            buf[so] = nxt[l].x; buf[so + 2] = nxt[l].y; buf[so + 4] = nxt[l].z; buf[so + 6] = nxt[l].w;
        }
        if (tid == 0) tb_lo[it & 1] = nbx;
        if (tid == 1) tb_hi[it & 1] = nbx;
        __syncthreads();
#pragma unroll
        for (int l = 0; l < LPT; l++) {
            const int nidx = s0 + (it + 1) * AP_TILE + l * AP_THREADS + tid;
            nxt[l] = (nidx < s1) ? src[nidx] : pad;
        }
        if (tid < 2 && it + 1 < ntiles) nbx = tbox[2 * (tile0 + it + 1) + tid];
        const float4 lo = tb_lo[it & 1], hi = tb_hi[it & 1];
        const bool far = lo.x > e[3] || lo.y > e[4] || lo.z > e[5] || hi.x < e[0] || hi.y < e[1] || hi.z < e[2];
        const ulonglong2* tp = reinterpret_cast<const ulonglong2*>(buf);
        if (far) tile_loop<TPT, false, EQM>(tp, xi, yi, zi, a2, AX, AY, AZ, PH);
        else tile_loop<TPT, true, EQM>(tp, xi, yi, zi, a2, AX, AY, AZ, PH);
    }
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * AP_THREADS + tid;
        if (t < nt) part[(size_t)blockIdx.y * nt + t] = make_float4(AX[k], AY[k], AZ[k], PH[k]);
    }
}

// Fixed-order sum over the source splits; removes the capped self term, applies the common mass (EQM), G and the sign
// of Phi.
__global__ void __launch_bounds__(256) k_gravity_reduce(const float4* __restrict__ part, int splits, const float4* __restrict__ posh,
                                                        const float4* __restrict__ posm, int t0, int t1, float G, float wscale,
                                                        int eqm, float4* __restrict__ grav, int32_t* __restrict__ npart,
                                                        int32_t* __restrict__ napprox) {
    int nt = t1 - t0;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    float4 s = part[t];
    for (int k = 1; k < splits; k++) {
        float4 q = part[(size_t)k * nt + t];
        s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
    }
    float h = posh[t0 + t].w;
    float self = (eqm ? 1.0f : posm[t0 + t].w) * rsqrt_approx(h * h);   // the self pair always sits in a capped tile
    float gs = G * wscale;
    grav[t0 + t] = make_float4(gs * s.x, gs * s.y, gs * s.z, -gs * (s.w - self));
    npart[t0 + t] = 0;
    napprox[t0 + t] = 0;
}

// FP32 FMA-pipe microbenchmark: 8 independent chains per thread.
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = fmaf(v[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += v[k];
    if (s == 12345.678f) out[0] = s;  // never true; keeps the chains alive
}

}  // namespace

int sph_launch_gravity_allpairs(sphb200_ctx* c) {
    int t0 = (int)c->t0;
    int t1 = (c->t1 < 0 || c->t1 > c->n) ? (int)c->n : (int)c->t1;
    int nt = t1 - t0;
    if (nt <= 0) return SPH_OK;
    int n = (int)c->gsrc_n;          // sources: every particle of the (global) sorted order
    const float4* src = c->gsrc;
    int tblocks = sph_div_up(nt, AP_THREADS * AP_TPT);
    // Source splits: the grid (target blocks x splits) should fill whole waves of resident CTAs (one 8-warp CTA per SM at 160
    // registers: asked from the occupancy calculator once) -- among the split counts the partial-sum buffer (gpart_entries
    // float4) and the tile granularity allow, take the one with the least idle tail, preferring >= 12 waves.
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        int a = 0, b = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_gravity_allpairs<AP_TPT, true>, AP_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_gravity_allpairs<AP_TPT, false>, AP_THREADS, 0);
        ctas_per_sm = std::max(1, std::min(a, b));
    }
    const int resident = c->sm_count * ctas_per_sm;
    int64_t smax = (int64_t)(c->gpart_entries / (size_t)nt);
    smax = std::min<int64_t>(smax, sph_div_up(n, AP_TILE));
    smax = std::max<int64_t>(std::min<int64_t>(smax, 65535), 1);
    int splits = 1, per = 0;
    auto efficiency = [&](int sp, int& ns, int& pr) {
        pr = sph_div_up(sph_div_up(n, sp), AP_TILE) * AP_TILE;       // sources per split: whole tiles
        ns = sph_div_up(n, pr);                                       // splits that are not empty
        const int64_t grid = (int64_t)tblocks * ns;
        const int64_t waves = (grid + resident - 1) / resident;
        double eff = (double)grid / (double)(waves * resident);
        if (waves < 12) eff *= 0.5 + 0.5 * (double)waves / 12.0;     // too few waves: the last one weighs more, tiles load less often
        return eff;
    };
    double best = -1.0;
    for (int sp = (int)smax; sp >= 1; sp--) { int ns, pr; best = std::max(best, efficiency(sp, ns, pr)); }
    for (int sp = 1; sp <= (int)smax; sp++) {                         // the fewest splits (least partial-sum traffic) within 0.5 % of the best
        int ns, pr;
        if (efficiency(sp, ns, pr) >= best - 0.005) { splits = ns; per = pr; break; }
    }
    k_tile_boxes<<<sph_div_up(n, AP_TILE), AP_TILE, 0, c->stream>>>(src, n, c->tbox);
    SPH_LAUNCH_CHECK(c);
    dim3 grid(tblocks, splits);
    if (c->equal_mass)
        k_gravity_allpairs<AP_TPT, true><<<grid, AP_THREADS, 0, c->stream>>>(src, n, per, c->tbox, c->posh[c->cur], t0, t1, c->gpart);
    else
        k_gravity_allpairs<AP_TPT, false><<<grid, AP_THREADS, 0, c->stream>>>(src, n, per, c->tbox, c->posh[c->cur], t0, t1, c->gpart);
    SPH_LAUNCH_CHECK(c);
    k_gravity_reduce<<<sph_div_up(nt, 256), 256, 0, c->stream>>>(c->gpart, splits, c->posh[c->cur], c->posm, t0, t1, c->p.G,
                                                                c->equal_mass ? c->common_mass : 1.0f, c->equal_mass ? 1 : 0, c->grav,
                                                                c->npart, c->napprox);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

int sph_fp32_peak(sphb200_ctx* c, double* tflops) {
    cudaEvent_t e0, e1;
    SPH_CK(c, cudaEventCreate(&e0));
    SPH_CK(c, cudaEventCreate(&e1));
    int blocks = c->sm_count * 8, iters = 4096;
    float* out = (float*)c->diag_d;
    k_fma_peak<<<blocks, 256, 0, c->stream>>>(out, 64, 1.0001f, 0.5f);  // warm-up
    SPH_LAUNCH_CHECK(c);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        SPH_CK(c, cudaEventRecord(e0, c->stream));
        k_fma_peak<<<blocks, 256, 0, c->stream>>>(out, iters, 1.0001f, 0.5f);
        SPH_LAUNCH_CHECK(c);
        SPH_CK(c, cudaEventRecord(e1, c->stream));
        SPH_CK(c, cudaEventSynchronize(e1));
        float ms = 0;
        SPH_CK(c, cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return SPH_OK;
}
