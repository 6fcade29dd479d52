// rowstream.cuh -- the access pattern shared by the kernels that sum over the finished neighbor rows (k_density,
// k_pressure_grad): one warp per 32 consecutive targets, rows consumed as one flat stream of OCTETS.
//
// The rows of the warp's targets are cut into octets (8 consecutive entries; the last octet of a row is padded) and the octets
// of all 32 rows form one stream that the warp consumes 8 at a time -- 8 lanes per octet, two octets per lane and iteration --
// so that every lane has a pair to evaluate whatever the spread of the row lengths (a fixed "16 lanes per target" mapping ran
// at 71 % lane use).  The row entries of the next iteration are fetched one iteration ahead (their DRAM latency hides behind
// the current iteration's gathers).
//
// Summation order is canonical: an octet is reduced by a 3-step butterfly over fixed row positions, and the octet sums of a
// row are added in ascending order by one lane in a second phase.  The result of a target therefore depends on its own row
// only, never on its 31 companions: single-GPU and sharded runs agree bit for bit without any alignment rule.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ctx.cuh"

constexpr int RS_WARPS = 8;     // warps per block
constexpr int RS_SLOTS = 320;   // octet sums per warp and round (32 rows x 46 neighbors = 200 octets; longer streams take rounds)

template <int NV>
struct RowStreamSmem {
    float4 tg[2][32];                 // two float4 records per target, filled by the caller
    int pre[33];                      // octet prefix of the 32 rows
    uint32_t om[RS_SLOTS + 16];       // octet of the round -> 16 * row | live entries << 9 | first entry (of the warp's rows) << 13
    float osum[RS_SLOTS][NV + 1];     // NV sums + a count per octet
};

// Runs the stream for one warp.  `clen` = length of the lane's own row (0: no row), `wrow` = row of the warp's first target (row q
// at wrow + q * kmax), `fallback` = any resident particle (what a padding lane evaluates; its result is dropped).
// pair(A, Bp, j, v, flag): one pair of the target with records (A, *Bp) with particle j -> NV values and a flag to count
// (the second record stays in shared memory until the callee reads it: a 128-bit shared load is four wavefronts of the L1 data
// pipe, which is what bounds these kernels).
// On return sum[0..NV) / count hold the lane's own target's totals.
template <int NV, bool COUNT, typename PairFn>
__device__ __forceinline__ void row_stream(RowStreamSmem<NV>& S, const uint32_t* __restrict__ wrow, uint32_t fallback, int kmax, int clen,
                                           PairFn pair, float (&sum)[NV], int& count) {
    constexpr unsigned FULLM = 0xffffffffu;
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    const int noct = (clen + 7) >> 3;
    int incl = noct;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULLM, incl, o);
        if (lane >= o) incl += v;
    }
    const int my_o0 = incl - noct;
    S.pre[lane] = my_o0;
    if (lane == 31) S.pre[32] = incl;
    __syncwarp();
    const char* tgw = reinterpret_cast<const char*>(&S.tg[0][0]);
#pragma unroll
    for (int k = 0; k < NV; k++) sum[k] = 0.f;
    count = 0;

    // entry `sub` of the octet described by m (beyond the row: the fallback particle)
    auto entry = [&](uint32_t m) {
        uint32_t j = fallback;
        const uint32_t* ep = wrow + ((m >> 13) + sub);
        if ((uint32_t)sub < ((m >> 9) & 0xfu)) SPH_DBG_IDX((m >> 13) + sub, 32 * kmax);
        asm("{ .reg .pred p; setp.lt.u32 p, %2, %3; @p ld.global.nc.u32 %0, [%1]; }" : "+r"(j) : "l"(ep), "r"((uint32_t)sub), "r"((m >> 9) & 0xfu));
        return j;
    };
    auto eval = [&](uint32_t m, uint32_t j, float (&v)[NV], bool& flag) {
        const float4 A = *reinterpret_cast<const float4*>(tgw + (m & 0x1f0u));
        const float4* Bp = reinterpret_cast<const float4*>(tgw + (m & 0x1f0u) + 32 * sizeof(float4));   // read by the callee, if at all
        const bool ok = sub < (int)((m >> 9) & 0xfu);
        pair(A, Bp, j, v, flag);
        flag = flag && ok;
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] = ok ? v[k] : 0.0f;
    };

    for (int r_lo = 0; r_lo < 32;) {
        // rows of this round: as many as fit the octet-sum slots (a row has at most 512 / 8 <= RS_SLOTS octets)
        const int o_lo = S.pre[r_lo];
        const bool fits = lane < r_lo || S.pre[lane + 1] - o_lo <= RS_SLOTS;
        const int r_hi = __popc(__ballot_sync(FULLM, fits));
        const int no = S.pre[r_hi] - o_lo;
        const bool mine = lane >= r_lo && lane < r_hi;
        if (mine)
            for (int u = 0; u < noct; u++) {
                SPH_DBG_IDX(my_o0 - o_lo + u, RS_SLOTS);
                S.om[my_o0 - o_lo + u] = (uint32_t)(lane * 16) | ((uint32_t)min(8, clen - 8 * u) << 9) | ((uint32_t)(lane * kmax + 8 * u) << 13);
            }
        if (lane < 16) S.om[no + lane] = 0u;   // padding of the last iteration and of the look-ahead: no live entries
        __syncwarp();
        uint32_t ma = S.om[grp], mb = S.om[4 + grp];
        uint32_t ja = entry(ma), jb = entry(mb);
        for (int ob = 0; ob < no; ob += 8) {
            const uint32_t na = S.om[ob + 8 + grp], nb = S.om[ob + 12 + grp];
            const uint32_t jna = entry(na), jnb = entry(nb);
            float va[NV], vb[NV];
            bool fa = false, fb = false;
            eval(ma, ja, va, fa);
            eval(mb, jb, vb, fb);
            ma = na; mb = nb; ja = jna; jb = jnb;
            unsigned ba = 0u, bb = 0u;
            if (COUNT) { ba = __ballot_sync(FULLM, fa); bb = __ballot_sync(FULLM, fb); }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1)
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    va[k] += __shfl_xor_sync(FULLM, va[k], o);
                    vb[k] += __shfl_xor_sync(FULLM, vb[k], o);
                }
            if (sub == 0) {   // the stream is padded to a multiple of 8 octets: slots beyond RS_SLOTS do not exist
                const int sa = ob + grp, sb = ob + 4 + grp;
                if (sa < RS_SLOTS) {
#pragma unroll
                    for (int k = 0; k < NV; k++) S.osum[sa][k] = va[k];
                    if (COUNT) S.osum[sa][NV] = __int_as_float(__popc((ba >> (8 * grp)) & 0xffu));
                }
                if (sb < RS_SLOTS) {
#pragma unroll
                    for (int k = 0; k < NV; k++) S.osum[sb][k] = vb[k];
                    if (COUNT) S.osum[sb][NV] = __int_as_float(__popc((bb >> (8 * grp)) & 0xffu));
                }
            }
        }
        __syncwarp();
        if (mine)
            for (int u = 0; u < noct; u++) {
                SPH_DBG_IDX(my_o0 - o_lo + u, RS_SLOTS);
                const float* v = S.osum[my_o0 - o_lo + u];
#pragma unroll
                for (int k = 0; k < NV; k++) sum[k] += v[k];
                if (COUNT) count += __float_as_int(v[NV]);
            }
        __syncwarp();
        r_lo = r_hi;
    }
}

__device__ __forceinline__ float rs_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rs_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
