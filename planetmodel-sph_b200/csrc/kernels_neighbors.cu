// kernels_neighbors.cu -- K1: neighbor lists, density and EOS.
//
// Replaces: the Unity.Physics broadphase pair stream + KernelSystem.FilterPairs / CalculateInteractionJob
// (UP/Collision/World/Broadphase.cs:275-351, A/Systems/KernelSystem.cs:234-335, 583-633), SplineKernel
// (A/Util/SplineKernel.cs:47-89) and DensityFieldSystem + the EOS (A/Systems/DensityFieldSystem.cs:38-56,
// A/Systems/PressureFieldSystem.cs:30-34).
//
// Three kernels (which one does the work is decided on the device from the grid's h_max, so the host never synchronises; the
// third is launched only while the host-side bound of h -- ctx.cuh h_bound -- cannot rule h_max >= 1e5 out):
//   k_cell_neighbors    h_max <  1e5 (every real run): cell-centric list build, one warp per pass of <= 32 targets of one
//                       cell, sqrt-free exact keep thresholds -- described at its definition below;
//   k_density           h_max <  1e5: density + EOS + own-support count from the finished rows, consumed as a flat octet stream
//                       (rowstream.cuh);
//   k_neighbors_density h_max >= 1e5 (W(r,h) can underflow, the threshold form is not valid): the warp-per-target kernel
//                       that evaluates the reference's literal Kernel(r,h) > 0 keep rule, lists + density + EOS in one
//                       pass.  One warp per target particle, three phases:
//     1. CELLS   lanes enumerate the (2S+1)^3 stencil cells around the target's cell, read (start, end, hmax) from the
//                cell table and cull every cell whose box is farther than 2*max(h_i, hmax_cell) from the target -- the
//                variable-h rule "r < 2 max(h_i,h_j)" without paying for the global h_max in every cell.
//     2. TEST    surviving cells are contiguous slot ranges of the Morton-sorted SoA; four 8-lane groups stream them
//                with coalesced float4 loads and apply the reference's exact fp32 predicate + keep rule.
//     3. SUM     survivors are queued in shared memory and evaluated 32 at a time (kernel values, density sum,
//                own-support count); list rows are written coalesced.
// Numerics contract: membership exact (non-contracted __f*_rn ops, IEEE sqrt), values fast (<= 1e-5 relative).
#include "ctx.cuh"
#include "rowstream.cuh"
#include <math.h>

namespace {

constexpr float kPI = 3.14159274f;  // Mathf.PI
constexpr float kInvPI = 0.318309886f;
constexpr unsigned FULL = 0xffffffffu;

constexpr float kHugeH = 1.0e5f;   // above it W(r,h) can underflow: the literal kernel decides membership (k_neighbors_density)

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
    const sph_GridParams* __restrict__ g, int t0, int t1, int rowbase, int kmax, float Keos, uint32_t* __restrict__ nlist,
    int32_t* __restrict__ ncount, int32_t* __restrict__ nown, float* __restrict__ rho, float* __restrict__ press,
    float* __restrict__ cvol, int32_t* __restrict__ err) {
    __shared__ uint2 cellq[K1_WARPS][K1_CELLQ];
    __shared__ Surv survq[K1_WARPS][K1_SURVQ];
    if (g->hmax < kHugeH) return;   // the cell-centric kernel below owns the normal case
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int bits = g->bits, S = g->stencil;
    const int shift = 3 * (10 - bits), dim = 1 << bits;
    const float fs = g->fine_scale, cw = (float)(1 << (10 - bits));  // cell width in fine units
    const float gx0 = g->min[0], gy0 = g->min[1], gz0 = g->min[2];
    const int nst = 2 * S + 1;
    const int grp = lane >> 3, sl = lane & 7;
    const unsigned lt = (1u << lane) - 1u;

    // persistent warps (the launch is a few blocks per SM: when the cell-centric kernel owns the step this one costs ~2 us)
    for (int warp = blockIdx.x * K1_WARPS + w; t0 + warp * K1_TPW < t1; warp += gridDim.x * K1_WARPS)
    for (int tt = 0; tt < K1_TPW; tt++) {
        const int t = t0 + warp * K1_TPW + tt;
        if (t >= t1) break;
        const float4 pi = posh[t];
        const float hi = pi.w, hi2 = __fmul_rn(hi, 2.0f), hinv_i = 1.0f / hi;
        // scaled (fine-grid) coordinates of the target, same ops as the key kernel
        const float sx = __fmul_rn(__fsub_rn(pi.x, gx0), fs), sy = __fmul_rn(__fsub_rn(pi.y, gy0), fs),
                    sz = __fmul_rn(__fsub_rn(pi.z, gz0), fs);
        const uint32_t ck = keys[t] >> shift;
        const int cx = (int)compact10(ck), cy = (int)compact10(ck >> 1), cz = (int)compact10(ck >> 2);
        uint32_t* row = nlist + (size_t)(t - rowbase) * kmax;
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
            if (count > kmax) { atomicMax(&err[ERR_NEIGHBOR_OVERFLOW], count); atomicMax(&err[ERR_OVERFLOW_EVER], count); }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// K1 (cell-centric): one warp per PASS = up to 32 consecutive targets of one cell (a cell holds ~11 particles when h is
// uniform, ~100 where h is much smaller than the typical h the cells are sized by).  The targets are staged in shared
// memory; the particles of the (2S+1)^3 stencil cells are the candidates: their cell ranges are flattened so that every
// lane fetches one candidate per batch, candidates that cannot reach the box of the pass's targets are dropped and the
// rest compacted into a shared-memory ring; every lane then owns one compacted candidate and tests it against each target
// in turn (LDS broadcast): 32 exact pair tests per ~30 instructions, no sqrt and no division in the test.
//
// Exactness without the sqrt: with s = max(h_i,h_j) the reference keeps the pair iff
//     d2 < ((s*s)*2)*2   and   ( fsqrt_rn(d2) < 2 h_i  or  fsqrt_rn(d2) < 2 h_j )      (SplineKernel.cs:47-53, :62)
// fsqrt_rn is monotone, so "fsqrt_rn(d2) < 2h" <=> d2 < t(h), t(h) = the smallest float whose rounded root reaches 2h;
// both bounds grow with h, hence keep <=> d2 < max(C_i, C_j) with C(h) = min(4 RN(h*h), t(h)) computed once per particle
// by the permute kernel (sph_keep_threshold, ctx.cuh) and carried in posc.w.  Valid below h = 1e5 (kHugeH).
//
// Cells of the outer stencil shells (S > 1: some h exceed the typical h the cells are sized by) are culled against the
// box of the pass's targets grown by the larger of the targets' and the cell's h_max.
// Rows are written by ballot rank: the order of a row depends only on the cell enumeration order, never on which warp or
// which rank built it (sharded runs stay bit-identical); density, EOS and the own-support count follow in k_density.
// ------------------------------------------------------------------------------------------------------------
constexpr int K3_WARPS = 8;
constexpr int K3_MINB = 3;    // resident blocks per SM (80 registers; 2 and 4 measured slower)
constexpr int K3_QCAP = 64;   // candidate-cell queue per warp
constexpr int K3_RING = 64;   // compacted-candidate ring per warp
constexpr int K3_MAXT = 32;   // targets per pass (cells are sized for ~11); consecutive sorted slots: a compact sub-box of the cell

template <bool EQM>
__global__ void __launch_bounds__(K3_WARPS * 32, K3_MINB) k_cell_neighbors(
    const float4* __restrict__ posc, const float4* __restrict__ posh, const float4* __restrict__ posm,
    const uint32_t* __restrict__ keys, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end,
    const uint32_t* __restrict__ cell_hmax, const sph_GridParams* __restrict__ g, int t0, int t1, int rowbase, int kmax, float Keos,
    uint32_t* __restrict__ nlist, int32_t* __restrict__ ncount, int32_t* __restrict__ nown, float* __restrict__ rho, float* __restrict__ press,
    float* __restrict__ cvol, int32_t* __restrict__ err, unsigned int* __restrict__ chunk_counter) {
    __shared__ float4 tgt[K3_WARPS][K3_MAXT];
    __shared__ uint32_t qstart[K3_WARPS][K3_QCAP];
    __shared__ uint32_t qpre[K3_WARPS][K3_QCAP + 1];
    __shared__ float4 ring4[K3_WARPS][K3_RING];
    __shared__ int ringj[K3_WARPS][K3_RING];
    __shared__ int cnts[K3_WARPS][K3_MAXT];
    if (!(g->hmax < kHugeH)) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int bits = g->bits, S = g->stencil, dim = 1 << bits;
    const int nst = 2 * S + 1, nst2 = nst * nst, nst3 = nst2 * nst;
    const int inv1 = ((1 << 20) + nst - 1) / nst, inv2 = ((1 << 20) + nst2 - 1) / nst2;
    const float fs = g->fine_scale, cw = (float)(1 << (10 - bits));
    const float g0[3] = {g->min[0], g->min[1], g->min[2]};
    float4* tg = tgt[w];
    uint32_t* qs_w = qstart[w];
    uint32_t* qp_w = qpre[w];
    float4* rb4 = ring4[w];
    int* rbj = ringj[w];
    int* cn_w = cnts[w];

    // Scheduling: the work items are the passes -- runs of <= 32 consecutive targets of one cell -- and a pass is owned by
    // the 32-slot block of the sorted array that holds its first target.  Warps pull 32-slot blocks from a counter and find
    // the pass heads in them from the sorted keys, so the unit of work is ~32 targets whatever the cell occupancy (cells
    // are sized by the typical h: in a region with much smaller h they hold ~100 particles), and empty cells cost nothing.
    const int shift = 3 * (10 - bits);
    const int ngrab = (t1 - t0 + 31) >> 5;
    while (true) {
        int grab = 0;
        if (lane == 0) grab = (int)atomicAdd(chunk_counter, 1u);
        grab = __shfl_sync(FULL, grab, 0);
        if (grab >= ngrab) break;
        const int slot = t0 + grab * 32 + lane;
        bool head = false;
        int my_e = 0;
        uint32_t my_ck = 0u;
        if (slot < t1) {
            my_ck = keys[slot] >> shift;
            const int cs = max((int)cell_start[my_ck], t0);
            my_e = min((int)cell_end[my_ck], t1);
            head = ((slot - cs) % K3_MAXT) == 0;
        }
        unsigned heads = __ballot_sync(FULL, head);
        while (heads) {
            const int src = __ffs(heads) - 1;
            heads &= heads - 1u;
            const int p0 = t0 + grab * 32 + src, e = __shfl_sync(FULL, my_e, src);
            const uint32_t ck = __shfl_sync(FULL, my_ck, src);
            const int cx = (int)compact10(ck), cy = (int)compact10(ck >> 1), cz = (int)compact10(ck >> 2);
            {
                const int nt = min(K3_MAXT, e - p0);   // targets of this pass (a cell far denser than its size: several passes)
                SPH_DBG_IDX(nt - 1, K3_MAXT);
                SPH_DBG_IDX(p0 - rowbase, 0x7fffffff);
                __syncwarp();
                int rh = 0, rt = 0;   // candidate ring head / tail (monotone counters)
                // stage the targets; box of their positions, their largest keep threshold and largest h
                float wlo[3] = {INFINITY, INFINITY, INFINITY}, whi[3] = {-INFINITY, -INFINITY, -INFINITY};
                float cmax = 0.f, hmax_t = 0.f;
                for (int i = lane; i < nt; i += 32) {
                    const float4 T = posc[p0 + i];
                    tg[i] = T;
                    cn_w[i] = 0;
                    wlo[0] = fminf(wlo[0], T.x); wlo[1] = fminf(wlo[1], T.y); wlo[2] = fminf(wlo[2], T.z);
                    whi[0] = fmaxf(whi[0], T.x); whi[1] = fmaxf(whi[1], T.y); whi[2] = fmaxf(whi[2], T.z);
                    cmax = fmaxf(cmax, T.w);
                    hmax_t = fmaxf(hmax_t, posh[p0 + i].w);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        wlo[k] = fminf(wlo[k], __shfl_xor_sync(FULL, wlo[k], o));
                        whi[k] = fmaxf(whi[k], __shfl_xor_sync(FULL, whi[k], o));
                    }
                    cmax = fmaxf(cmax, __shfl_xor_sync(FULL, cmax, o));
                    hmax_t = fmaxf(hmax_t, __shfl_xor_sync(FULL, hmax_t, o));
                }
                uint32_t* rowp = nlist + (size_t)(p0 - rowbase) * kmax;
                int qn = 0, qtotal = 0;

                // 32 compacted candidates (ring entries rh .. rh+m-1) against every target of the pass; row fill counts live
                // in shared memory (warp-uniform reads, lane 0 writes)
                auto test_batch = [&](int m) {
                    const int er = (rh + lane) & (K3_RING - 1);
                    float4 c = rb4[er];
                    const int j = rbj[er];
                    if (lane >= m) c.x = INFINITY;       // idle lane: d2 = inf, never kept
                    const int jrel = j - p0;
                    uint32_t* rowt = rowp;
#pragma unroll 2
                    for (int tt = 0; tt < nt; tt++, rowt += kmax) {
                        const float4 Tt = tg[tt];
                        const int base = cn_w[tt];
                        const float dx = __fsub_rn(Tt.x, c.x), dy = __fsub_rn(Tt.y, c.y), dz = __fsub_rn(Tt.z, c.z);
                        const float d2 = dot3_rn(dx, dy, dz);
                        const bool keep = d2 < fmaxf(Tt.w, c.w) && jrel != tt;
                        const unsigned kb = __ballot_sync(FULL, keep);
                        const int slot = base + __popc(kb & lt);
                        if (keep) { SPH_DBG_IDX(slot, 1 << 24); SPH_DBG_IDX(j, t1 + (1 << 28)); }
                        if (keep && slot < kmax) rowt[slot] = (uint32_t)j;   // beyond max_neighbors: counted, not stored
                        if (lane == 0) cn_w[tt] = base + __popc(kb);
                    }
                    rh += m;
                };
                // streams the queued cells (flattened: one candidate per lane per batch, next batch's loads in flight);
                // a candidate enters the ring only if it can reach the box of the targets:
                // d2(i,j) < max(C_i, C_j) for some target i needs dist^2(p_j, box) (1 - 1e-5) < max(C_max, C_j)
                auto flush = [&]() {
                    SPH_DBG_IDX(qn, K3_QCAP + 1);
                    if (lane == 0) qp_w[qn] = (uint32_t)qtotal;
                    __syncwarp();
                    int q0 = 0;   // queue entry holding flattened position f0 of the batch being fetched
                    auto fetch = [&](int f0, int& j, float4& c) {
                        // entries q0+1 .. q0+32 that start within positions f0+1 .. f0+32 (lengths are >= 1, so there are
                        // at most 32): bit (pos-1) per start; lane L sits in entry q0 + #starts at positions <= L
                        const int qi = q0 + 1 + lane;
                        const int pos = qi < qn ? (int)qp_w[qi] - f0 : 64;
                        const unsigned starts = __reduce_or_sync(FULL, pos <= 32 ? 1u << (pos - 1) : 0u);
                        const int q = q0 + __popc(starts & lt);
                        const int f = f0 + lane;
                        if (f < qtotal) SPH_DBG_IDX(q, qn);
                        j = f < qtotal ? (int)(qs_w[q] + ((uint32_t)f - qp_w[q])) : 0;
                        c = posc[j];
                        q0 += __popc(starts);
                    };
                    int j; float4 c;
                    fetch(0, j, c);
                    for (int f0 = 0; f0 < qtotal; f0 += 32) {
                        int jn = 0; float4 cn = c;
                        if (f0 + 32 < qtotal) fetch(f0 + 32, jn, cn);
                        const float gx = fmaxf(fmaxf(wlo[0] - c.x, c.x - whi[0]), 0.f), gy = fmaxf(fmaxf(wlo[1] - c.y, c.y - whi[1]), 0.f),
                                    gz = fmaxf(fmaxf(wlo[2] - c.z, c.z - whi[2]), 0.f);
                        const bool pre = f0 + lane < qtotal && (gx * gx + gy * gy + gz * gz) * 0.99999f < fmaxf(cmax, c.w);
                        const unsigned pb = __ballot_sync(FULL, pre);
                        if (pre) {
                            const int er = (rt + __popc(pb & lt)) & (K3_RING - 1);
                            rb4[er] = c;
                            rbj[er] = j;
                        }
                        rt += __popc(pb);
                        __syncwarp();
                        if (rt - rh >= 32) { test_batch(32); __syncwarp(); }
                        j = jn; c = cn;
                    }
                    qn = 0; qtotal = 0;
                };

                // ---- stencil cells; the outer shells (S > 1) are culled against the box of the targets
                float plo[3], phi[3], reach_t = 0.f;
                if (S > 1) {
#pragma unroll
                    for (int k = 0; k < 3; k++) {   // same ops as the key kernel: monotone, so the box maps to a box
                        plo[k] = __fmul_rn(__fsub_rn(wlo[k], g0[k]), fs);
                        phi[k] = __fmul_rn(__fsub_rn(whi[k], g0[k]), fs);
                    }
                    reach_t = 2.0f * hmax_t * fs * 1.001f + 0.01f;
                }
                for (int base = 0; base < nst3; base += 32) {
                    const int k = base + lane;
                    // k < 729, divisors <= 81: (k * ceil(2^20/d)) >> 20 == k / d
                    const int oz = (k * inv2) >> 20, rem = k - oz * nst2, oy = (rem * inv1) >> 20, ox = rem - oy * nst;
                    const int nx = cx + ox - S, ny = cy + oy - S, nz = cz + oz - S;
                    bool ok = k < nst3 && nx >= 0 && ny >= 0 && nz >= 0 && nx < dim && ny < dim && nz < dim;
                    uint32_t a = 0, b = 0;
                    if (ok) {
                        const uint32_t nk = expand10((uint32_t)nx) | (expand10((uint32_t)ny) << 1) | (expand10((uint32_t)nz) << 2);
                        a = cell_start[nk]; b = cell_end[nk];
                        ok = b > a;
                        if (ok && S > 1 && (abs(ox - S) > 1 || abs(oy - S) > 1 || abs(oz - S) > 1)) {
                            const float blo[3] = {(float)nx * cw, (float)ny * cw, (float)nz * cw};
                            float gp2 = 0.f;
#pragma unroll
                            for (int d = 0; d < 3; d++) {
                                const float gp = fmaxf(fmaxf(blo[d] - phi[d], plo[d] - (blo[d] + cw)), 0.f);
                                gp2 += gp * gp;
                            }
                            // conservative reach of the larger of the targets' and the cell's h: +0.1% and +0.01 fine units
                            // cover every rounding of the scaled coordinates
                            const float rc = fmaxf(reach_t, 2.0f * __uint_as_float(cell_hmax[nk]) * fs * 1.001f + 0.01f);
                            ok = gp2 < rc * rc;
                        }
                    }
                    const unsigned bal = __ballot_sync(FULL, ok);
                    if (bal == 0u) continue;
                    int incl = ok ? (int)(b - a) : 0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (ok) {
                        const int pos = qn + __popc(bal & lt);
                        SPH_DBG_IDX(pos, K3_QCAP);
                        qs_w[pos] = a;
                        qp_w[pos] = (uint32_t)(qtotal + incl - (int)(b - a));
                    }
                    qn += __popc(bal);
                    qtotal += __shfl_sync(FULL, incl, 31);
                    if (qn > K3_QCAP - 32) flush();
                }
                if (qn > 0) flush();
                if (rt - rh > 0) test_batch(rt - rh);
                __syncwarp();

                // counts now; density + EOS + own-support count follow in k_density (rows are complete after this kernel)
                for (int i = lane; i < nt; i += 32) {
                    const int cnt = cn_w[i];
                    ncount[p0 + i] = cnt;
                    if (cnt > kmax) { atomicMax(&err[ERR_NEIGHBOR_OVERFLOW], cnt); atomicMax(&err[ERR_OVERFLOW_EVER], cnt); }
                }
            }
        }
    }
}


// K1b: density + EOS + own-support count from the finished rows (DensityFieldSystem.cs:38-56, PressureFieldSystem.cs:30-34).
// One warp per 32 consecutive targets, rows consumed as a flat stream of octets with canonical summation order (rowstream.cuh).
// Own-support membership: fsqrt_rn(d2) < 2 h_i with the exact d2 (SplineKernel.cs:62) <=> d2 < t(h_i), t(h) = the smallest
// float whose rounded root reaches 2h (sph_own_threshold): no IEEE sqrt per pair.
// (pi h^3) W(q) = 0.25 (2-q)+^3 - (1-q)+^3  (SplineKernel.cs:55-89, branch-free)
__device__ __forceinline__ float w_shape(float q) {
    const float t1 = fmaxf(2.0f - q, 0.0f), t2 = fmaxf(1.0f - q, 0.0f);
    return fmaf(0.25f * t1, t1 * t1, -(t2 * t2) * t2);
}

template <bool EQM>
__global__ void __launch_bounds__(RS_WARPS * 32, 4) k_density(const float4* __restrict__ posh, const float4* __restrict__ posm,
                                                              const float4* __restrict__ posc, const uint32_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end,
                                                              const uint32_t* __restrict__ nlist, const int32_t* __restrict__ ncount,
                                                              const sph_GridParams* __restrict__ g, int t0, int t1, int rowbase, int kmax, float Keos,
                                                              int32_t* __restrict__ nown, float* __restrict__ rho, float* __restrict__ press,
                                                              float* __restrict__ cvol) {
    __shared__ RowStreamSmem<1> smem[RS_WARPS];   // tg[0]: x, y, z, 1/h   tg[1]: t(h), 1/(pi h^3), -, -
    if (!(g->hmax < kHugeH)) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int base = t0 + (blockIdx.x * RS_WARPS + w) * 32;
    if (base >= t1) return;
    RowStreamSmem<1>& S = smem[w];
    const int t = base + lane;
    const bool live = t < t1;
    float4 pi = make_float4(0.f, 0.f, 0.f, 1.f);
    int cnt = 0;
    if (live) { pi = posh[t]; cnt = ncount[t]; }
    const bool ovf = cnt > kmax;   // truncated row (an error state the caller is told about): re-scanned from the stencil below
    const float hinv_i = 1.0f / pi.w;
    S.tg[0][lane] = make_float4(pi.x, pi.y, pi.z, hinv_i);
    S.tg[1][lane] = make_float4(sph_own_threshold(pi.w), hinv_i * hinv_i * hinv_i * kInvPI, 0.f, 0.f);
    float acc[1];
    int own;
    row_stream<1, true>(S, nlist + (size_t)(base - rowbase) * kmax, (uint32_t)base, kmax, ovf ? 0 : cnt,
        [&](const float4 A, const float4* Bp, uint32_t j, float (&v)[1], bool& in_i) {
            const float2 B = *reinterpret_cast<const float2*>(Bp);   // t(h_i), 1/(pi h_i^3)
            const float4 pj = posh[j];
            const float dx = __fsub_rn(A.x, pj.x), dy = __fsub_rn(A.y, pj.y), dz = __fsub_rn(A.z, pj.z);
            const float d2 = dot3_rn(dx, dy, dz);
            in_i = d2 < B.x;
            const float r = d2 * rs_rsqrt(fmaxf(d2, 1.0e-37f));
            const float hinv_j = rs_rcp(pj.w);
            const float cw_j = hinv_j * hinv_j * hinv_j * kInvPI;
            float wsum = fmaf(B.y, w_shape(r * A.w), cw_j * w_shape(r * hinv_j));
            if (!EQM) wsum *= __ldg(&posm[j].w);
            v[0] = wsum;
        }, acc, own);
    float rsum = acc[0];

    // Row overflow: the list is truncated, but the density and the own-support count stay complete -- the whole warp
    // re-scans the stencil of such a target with the keep rule of k_cell_neighbors (lane-strided partial sums, butterfly).
    unsigned ovm = __ballot_sync(FULL, ovf);
    while (ovm) {
        const int src = __ffs(ovm) - 1;
        ovm &= ovm - 1u;
        const int tq = base + src;
        const float4 A = S.tg[0][src];
        const float4 C = S.tg[1][src];
        const int bits = g->bits, S = g->stencil, dim = 1 << bits;
        const uint32_t ck = keys[tq] >> (3 * (10 - bits));
        const int cx = (int)compact10(ck), cy = (int)compact10(ck >> 1), cz = (int)compact10(ck >> 2);
        const float c_t = posc[tq].w;
        float s = 0.f;
        int no = 0;
        for (int oz = -S; oz <= S; oz++)
            for (int oy = -S; oy <= S; oy++)
                for (int ox = -S; ox <= S; ox++) {
                    const int nx = cx + ox, ny = cy + oy, nz = cz + oz;
                    if (nx < 0 || ny < 0 || nz < 0 || nx >= dim || ny >= dim || nz >= dim) continue;
                    const uint32_t nk = expand10((uint32_t)nx) | (expand10((uint32_t)ny) << 1) | (expand10((uint32_t)nz) << 2);
                    for (uint32_t j = cell_start[nk] + lane; j < cell_end[nk]; j += 32) {
                        const float4 cj = posc[j];
                        const float dx = __fsub_rn(A.x, cj.x), dy = __fsub_rn(A.y, cj.y), dz = __fsub_rn(A.z, cj.z);
                        const float d2 = dot3_rn(dx, dy, dz);
                        if (d2 < fmaxf(c_t, cj.w) && j != (uint32_t)tq) {
                            const float hj = posh[j].w;
                            const float mj = EQM ? 1.0f : posm[j].w;
                            no += d2 < C.x ? 1 : 0;
                            const float r = d2 * rs_rsqrt(fmaxf(d2, 1.0e-37f));
                            const float hinv_j = rs_rcp(hj);
                            const float wsum = fmaf(C.y, w_shape(r * A.w), hinv_j * hinv_j * hinv_j * kInvPI * w_shape(r * hinv_j));
                            s += EQM ? wsum : mj * wsum;
                        }
                    }
                }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(FULL, s, o);
            no += __shfl_xor_sync(FULL, no, o);
        }
        if (lane == src) { rsum = s; own = no; }
    }

    if (live) {
        // self term m_i * Kernel(0,h_i) (DensityFieldSystem.cs:45), exact: 1/(pi h^3); rsum holds sum (W_i + W_j), halved here
        const float mi = posm[t].w;
        const float w0 = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(__fmul_rn(kPI, pi.w), pi.w), pi.w));
        rsum *= EQM ? 0.5f * mi : 0.5f;
        const float d = __fadd_rn(__fmul_rn(mi, w0), rsum);
        const float P = __fmul_rn(__fmul_rn(Keos, d), d);  // PressureFieldSystem.cs:31-33
        rho[t] = d;
        press[t] = P;
        cvol[t] = __fmul_rn(__fdiv_rn(mi, d), P);          // m_j / rho_j * P_j (PressureFieldSystem.cs:65)
        nown[t] = own;
    }
}

}  // namespace

int sph_launch_neighbors_density(sphb200_ctx* c) {
    int t0 = (int)c->t0;
    int t1 = (c->t1 < 0 || c->t1 > c->n) ? (int)c->n : (int)c->t1;
    if (t0 > t1) t0 = t1;
    int nt = t1 - t0;
    if (nt <= 0) return SPH_OK;
    // cell-centric kernel (h_max < 1e5): persistent warps pull 32-cell chunks from a counter
    SPH_CK(c, cudaMemsetAsync(c->chunk_counter, 0, sizeof(unsigned int), c->stream));
    // the overflow flag describes the lists this pass builds: a later pass whose rows fit clears it
    SPH_CK(c, cudaMemsetAsync(c->err_d + ERR_NEIGHBOR_OVERFLOW, 0, sizeof(int32_t), c->stream));
#define K3_LAUNCH(E) k_cell_neighbors<E><<<c->sm_count * K3_MINB, K3_WARPS * 32, 0, c->stream>>>(                                        \
        c->posc, c->posh[c->cur], c->posm, c->skeys, c->cell_start, c->cell_end, c->cell_hmax, c->grid_d, t0, t1, (int)c->row_base, c->p.max_neighbors, c->p.K, \
        c->nlist, c->ncount, c->nown, c->rho, c->press, c->cvol, c->err_d, c->chunk_counter)
    if (c->equal_mass) K3_LAUNCH(true); else K3_LAUNCH(false);
#undef K3_LAUNCH
    SPH_LAUNCH_CHECK(c);
    {
        int tpb = RS_WARPS * 32;   // targets per block: 32 per warp
        if (c->equal_mass)
            k_density<true><<<sph_div_up(nt, tpb), 256, 0, c->stream>>>(c->posh[c->cur], c->posm, c->posc, c->skeys, c->cell_start,
                                                                        c->cell_end, c->nlist, c->ncount, c->grid_d, t0, t1,
                                                                        (int)c->row_base, c->p.max_neighbors, c->p.K, c->nown, c->rho, c->press, c->cvol);
        else
            k_density<false><<<sph_div_up(nt, tpb), 256, 0, c->stream>>>(c->posh[c->cur], c->posm, c->posc, c->skeys, c->cell_start,
                                                                         c->cell_end, c->nlist, c->ncount, c->grid_d, t0, t1,
                                                                         (int)c->row_base, c->p.max_neighbors, c->p.K, c->nown, c->rho, c->press, c->cvol);
    }
    SPH_LAUNCH_CHECK(c);
    // literal-kernel variant: needed only when h_max >= 1e5 (decided on the device: no host sync); not launched at all while the
    // host-side bound of h (ctx.cuh h_bound) rules that out
    if (c->h_bound < kHugeH) return SPH_OK;
    int per_block = K1_WARPS * K1_TPW;
    k_neighbors_density<<<min(sph_div_up(nt, per_block), c->sm_count * 4), K1_WARPS * 32, 0, c->stream>>>(
        c->posh[c->cur], c->posm, c->skeys, c->cell_start, c->cell_end, c->cell_hmax, c->grid_d, t0, t1, (int)c->row_base, c->p.max_neighbors,
        c->p.K, c->nlist, c->ncount, c->nown, c->rho, c->press, c->cvol, c->err_d);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}
