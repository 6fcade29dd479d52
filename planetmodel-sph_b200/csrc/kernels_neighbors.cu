// kernels_neighbors.cu -- K1: neighbor lists + density + EOS in one pass (warp-cooperative gather).
//
// Replaces: the Unity.Physics broadphase pair stream + KernelSystem.FilterPairs / CalculateInteractionJob
// (UP/Collision/World/Broadphase.cs:275-351, A/Systems/KernelSystem.cs:234-335, 583-633), SplineKernel
// (A/Util/SplineKernel.cs:47-89) and DensityFieldSystem + the EOS (A/Systems/DensityFieldSystem.cs:38-56,
// A/Systems/PressureFieldSystem.cs:30-34).
//
// One warp per target particle, three phases, each with all 32 lanes busy:
//   1. CELLS   lanes enumerate the (2S+1)^3 stencil cells around the target's cell, read (start, end, hmax) from the
//              cell table and cull every cell whose box is farther than 2*max(h_i, hmax_cell) from the target -- the
//              variable-h rule "r < 2 max(h_i,h_j)" without paying for the global h_max in every cell.
//   2. TEST    surviving cells are contiguous slot ranges of the Morton-sorted SoA; four 8-lane groups stream them
//              with coalesced float4 loads and apply the reference's exact fp32 predicate + keep rule.
//   3. SUM     survivors are queued in shared memory and evaluated 32 at a time (kernel values, density sum,
//              own-support count); list rows are written coalesced.
// Numerics contract: membership exact (non-contracted __f*_rn ops, IEEE sqrt), values fast (<= 1e-5 relative).
#include "ctx.cuh"
#include <math.h>

namespace {

constexpr float kPI = 3.14159274f;  // Mathf.PI
constexpr float kInvPI = 0.318309886f;
constexpr unsigned FULL = 0xffffffffu;

constexpr int K1_WARPS = 8;
constexpr int K1_TPW = 4;     // consecutive targets per warp (L1 reuse of the neighbor cells)
constexpr int K1_CELLQ = 96;  // relevant-cell queue per warp
constexpr int K1_SURVQ = 64;  // survivor ring per warp

__device__ __forceinline__ float kernel_exact(float distance, float size) {  // SplineKernel.cs:55-89, op for op
    if (distance >= __fmul_rn(size, 2.0f)) return 0.0f;
    float q = __fdiv_rn(distance, size);
    float pi_h_cube = __fmul_rn(__fmul_rn(__fmul_rn(kPI, size), size), size);
    if (distance < size) {
        float q2 = __fmul_rn(q, q);
        float num = __fadd_rn(__fsub_rn(1.0f, __fmul_rn(1.5f, q2)), __fmul_rn(0.75f, __fmul_rn(q2, q)));
        return __fdiv_rn(num, pi_h_cube);
    }
    float t = __fsub_rn(2.0f, q);
    return __fdiv_rn(__fmul_rn(__fmul_rn(t, t), t), __fmul_rn(4.0f, pi_h_cube));
}

// M4 spline via 4(1 - 1.5q^2 + 0.75q^3) = (2-q)^3 - 4(1-q)^3 (branch-free), norm = 1/(pi h^3)
__device__ __forceinline__ float w_fast(float r, float hinv) {
    float q = r * hinv;
    float t1 = fmaxf(2.0f - q, 0.0f), t2 = fmaxf(1.0f - q, 0.0f);
    float c = hinv * hinv * hinv * (0.25f * kInvPI);
    return c * (t1 * t1 * t1 - 4.0f * (t2 * t2 * t2));
}

struct Surv { uint32_t j; float r; float hj; };  // r carries "inside i's own support" in its mantissa LSB

__global__ void __launch_bounds__(K1_WARPS * 32) k_neighbors_density(
    const float4* __restrict__ posh, const float4* __restrict__ posm, const uint32_t* __restrict__ keys,
    const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end, const uint32_t* __restrict__ cell_hmax,
    const sph_GridParams* __restrict__ g, int t0, int t1, int kmax, float Keos, uint32_t* __restrict__ nlist,
    int32_t* __restrict__ ncount, int32_t* __restrict__ nown, float* __restrict__ rho, float* __restrict__ press,
    float* __restrict__ cvol, int32_t* __restrict__ err) {
    __shared__ uint2 cellq[K1_WARPS][K1_CELLQ];
    __shared__ Surv survq[K1_WARPS][K1_SURVQ];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int warp = blockIdx.x * K1_WARPS + w;
    const int bits = g->bits, S = g->stencil;
    const int shift = 3 * (10 - bits), dim = 1 << bits;
    const float fs = g->fine_scale, cw = (float)(1 << (10 - bits));  // cell width in fine units
    const float gx0 = g->min[0], gy0 = g->min[1], gz0 = g->min[2];
    const int nst = 2 * S + 1;
    const int grp = lane >> 3, sl = lane & 7;
    const unsigned lt = (1u << lane) - 1u;

    for (int tt = 0; tt < K1_TPW; tt++) {
        const int t = t0 + warp * K1_TPW + tt;
        if (t >= t1) return;
        const float4 pi = posh[t];
        const float hi = pi.w, hi2 = __fmul_rn(hi, 2.0f), hinv_i = 1.0f / hi;
        // scaled (fine-grid) coordinates of the target, same ops as the key kernel
        const float sx = __fmul_rn(__fsub_rn(pi.x, gx0), fs), sy = __fmul_rn(__fsub_rn(pi.y, gy0), fs),
                    sz = __fmul_rn(__fsub_rn(pi.z, gz0), fs);
        const uint32_t ck = keys[t] >> shift;
        const int cx = (int)compact10(ck), cy = (int)compact10(ck >> 1), cz = (int)compact10(ck >> 2);
        uint32_t* row = nlist + (size_t)t * kmax;
        float rho_l = 0.f;
        int own_l = 0, count = 0;
        int qn = 0;            // relevant cells queued
        int sh = 0, st = 0;    // survivor ring head / tail (monotone counters)

        // evaluates 32 queued survivors (or the last `m` of them): all lanes busy
        auto drain = [&](int m) {
            if (lane < m) {
                Surv s = survq[w][(sh + lane) & (K1_SURVQ - 1)];
                bool in_i = __float_as_uint(s.r) & 1u;
                float wsym = 0.5f * (w_fast(s.r, hinv_i) + w_fast(s.r, __fdividef(1.0f, s.hj)));
                rho_l = fmaf(posm[s.j].w, wsym, rho_l);
                own_l += in_i ? 1 : 0;
                int slot = count + lane;
                if (slot < kmax) row[slot] = s.j;
            }
            count += m;
            sh += m;
        };
        // streams the queued cells: 4 groups of 8 lanes, one cell per group per round
        auto scan_cells = [&]() {
            for (int r0 = 0; r0 < qn; r0 += 4) {
                int c = r0 + grp;
                uint2 se = c < qn ? cellq[w][c] : make_uint2(0u, 0u);
                uint32_t j = se.x + sl;
                while (__any_sync(FULL, j < se.y)) {
                    bool keep = false, in_i = false;
                    float r = 0.f, hj = 1.f;
                    if (j < se.y && j != (uint32_t)t) {
                        float4 pj = posh[j];
                        hj = pj.w;
                        float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y), dz = __fsub_rn(pi.z, pj.z);
                        float d2 = dot3_rn(dx, dy, dz);
                        float szm = fmaxf(hi, hj);
                        // SplineKernel.Interacts (SplineKernel.cs:47-53): d2 < size*size*Kappa*Kappa
                        if (d2 < __fmul_rn(__fmul_rn(__fmul_rn(szm, szm), 2.0f), 2.0f)) {
                            r = __fsqrt_rn(d2);
                            if (szm < 1.0e5f) {
                                // keep rule KernelSymmetric.w > 0 (KernelSystem.cs:269,283): below h = 1e5 nothing can
                                // underflow, so W(r,h) > 0 <=> r < 2h (SplineKernel.cs:62)
                                in_i = r < hi2;
                                keep = in_i || (r < __fmul_rn(hj, 2.0f));
                            } else {
                                float wi = kernel_exact(r, hi), wj = kernel_exact(r, hj);
                                in_i = wi > 0.0f;
                                keep = __fmul_rn(__fadd_rn(wi, wj), 0.5f) > 0.0f;
                            }
                        }
                    }
                    unsigned bal = __ballot_sync(FULL, keep);
                    if (keep) {
                        Surv s;
                        s.j = j;
                        s.r = __uint_as_float((__float_as_uint(r) & ~1u) | (in_i ? 1u : 0u));
                        s.hj = hj;
                        survq[w][(st + __popc(bal & lt)) & (K1_SURVQ - 1)] = s;
                    }
                    st += __popc(bal);
                    __syncwarp();
                    if (st - sh >= 32) { drain(32); __syncwarp(); }
                    j += 8;
                }
            }
            qn = 0;
        };

        // ---- phase 1: cull the stencil cells.  Lanes own (x,y) columns of the stencil (their Morton xy-part and xy-gap
        // are computed once per target); the z planes are walked by the whole warp, so nothing is divided per cell.
        const float reach_i = 2.0f * hi * fs * 1.001f + 0.01f;
        for (int cbase = 0; cbase < nst * nst; cbase += 32) {
            const int col = cbase + lane;
            const int oy = col / nst, ox = col - oy * nst;
            const int nx = cx + ox - S, ny = cy + oy - S;
            const bool col_ok = col < nst * nst && nx >= 0 && ny >= 0 && nx < dim && ny < dim;
            const uint32_t kxy = expand10((uint32_t)nx) | (expand10((uint32_t)ny) << 1);
            // gap between the target and the cell box, in fine units (boxes are exact there)
            const float ax = fmaxf(fmaxf((float)nx * cw - sx, sx - (float)(nx + 1) * cw), 0.f);
            const float ay = fmaxf(fmaxf((float)ny * cw - sy, sy - (float)(ny + 1) * cw), 0.f);
            const float axy2 = ax * ax + ay * ay;
            for (int oz = 0; oz < nst; oz++) {
                const int nz = cz + oz - S;
                if (nz < 0 || nz >= dim) continue;   // warp-uniform
                const float az = fmaxf(fmaxf((float)nz * cw - sz, sz - (float)(nz + 1) * cw), 0.f);
                const float d2 = axy2 + az * az;
                bool rel = false;
                uint32_t s = 0, e = 0;
                // cheap pre-cull with the global h_max, then the per-cell h_max
                if (col_ok) {
                    const uint32_t nk = kxy | (expand10((uint32_t)nz) << 2);
                    s = cell_start[nk];
                    e = cell_end[nk];
                    if (e > s) {
                        float reach = reach_i;
                        if (d2 >= reach * reach) {
                            // conservative reach: +0.1% and +0.01 fine units cover every rounding of the scaled coordinates
                            float hm = __uint_as_float(cell_hmax[nk]);
                            reach = 2.0f * fmaxf(hi, hm) * fs * 1.001f + 0.01f;
                        }
                        rel = d2 < reach * reach;
                    }
                }
                unsigned bal = __ballot_sync(FULL, rel);
                if (rel) cellq[w][qn + __popc(bal & lt)] = make_uint2(s, e);
                qn += __popc(bal);
                __syncwarp();
                if (qn > K1_CELLQ - 32) scan_cells();  // ---- phase 2 (queue nearly full)
            }
        }
        scan_cells();                               // ---- phase 2 (rest)
        if (st - sh > 0) drain(st - sh);            // ---- phase 3 tail
        __syncwarp();

        float rsum = rho_l;
        int own = own_l;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rsum += __shfl_xor_sync(FULL, rsum, o);
            own += __shfl_xor_sync(FULL, own, o);
        }
        if (lane == 0) {
            // self term m_i * Kernel(0,h_i) (DensityFieldSystem.cs:45), exact: 1/(pi h^3)
            float mi = posm[t].w;
            float w0 = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(__fmul_rn(kPI, hi), hi), hi));
            float d = __fadd_rn(__fmul_rn(mi, w0), rsum);
            float P = __fmul_rn(__fmul_rn(Keos, d), d);  // PressureFieldSystem.cs:31-33
            rho[t] = d;
            press[t] = P;
            cvol[t] = __fmul_rn(__fdiv_rn(mi, d), P);    // m_j / rho_j * P_j (PressureFieldSystem.cs:65)
            ncount[t] = count;
            nown[t] = own;
            if (count > kmax) atomicMax(&err[ERR_NEIGHBOR_OVERFLOW], count);
        }
    }
}

}  // namespace

int sph_launch_neighbors_density(sphb200_ctx* c) {
    int t0 = (int)c->t0;
    int t1 = (c->t1 < 0 || c->t1 > c->n) ? (int)c->n : (int)c->t1;
    if (t0 > t1) t0 = t1;
    int nt = t1 - t0;
    if (nt <= 0) return SPH_OK;
    int per_block = K1_WARPS * K1_TPW;
    k_neighbors_density<<<sph_div_up(nt, per_block), K1_WARPS * 32, 0, c->stream>>>(
        c->posh[c->cur], c->posm, c->keys[1], c->cell_start, c->cell_end, c->cell_hmax, c->grid_d, t0, t1, c->p.max_neighbors,
        c->p.K, c->nlist, c->ncount, c->nown, c->rho, c->press, c->cvol, c->err_d);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}
