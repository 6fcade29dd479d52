// kernels_grid.cu -- smoothing-length controller, global bounds, Morton keys, radix sort, SoA permute, cell table.
//
// Replaces (behaviour, not code): ParticleSmoothingSystem (A/Systems/ParticleSmoothingSystem.cs:21-86),
// Broadphase AABB prep + BVH build (UP/Collision/World/Broadphase.cs:725-782,
// UP/Collision/Geometry/BoundingVolumeHierarchyBuilder.cs:416-467) and KernelSystem's flatten + two counting
// sorts (A/Systems/KernelSystem.cs:411-464, 539-568, 638-663).  Key arithmetic is specified by
// oracle/sph_oracle.cpp (orc_grid_params / orc_morton_keys) and must match it bit-for-bit.
#include "ctx.cuh"
#include <math.h>

namespace {

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// h <- (h*0.5)*(1 + rr[n_own])  (ParticleSmoothingSystem.cs:46-59; rr = correctly rounded pow(target/n, 1/3f),
// tabulated on the host so that it is bit-identical to the oracle), fused with the min/max/hmax reduction.
__global__ void __launch_bounds__(256) k_smoothing_bounds(float4* __restrict__ posh, const int32_t* __restrict__ nown,
                                                          const float* __restrict__ rr, int n, int update_h,
                                                          uint32_t* __restrict__ bounds) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float hmax = 0.f;
    unsigned long long bitsum = 0ull;  // sum of the fp32 bit patterns of h: integer, hence order-independent
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = posh[i];
        if (update_h) {
            int c = nown[i];
            if (c != 0) {
                c = min(c, SPH_RR_TABLE - 1);
                p.w = __fmul_rn(__fmul_rn(p.w, 0.5f), __fadd_rn(1.0f, rr[c]));
                posh[i].w = p.w;
            }
        }
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
        hmax = fmaxf(hmax, p.w);
        bitsum += __float_as_uint(p.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bitsum += __shfl_xor_sync(0xffffffffu, bitsum, o);
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)(bounds + 8), bitsum);
    __shared__ float s[7][8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float v[7] = {warp_min(lo[0]), warp_min(lo[1]), warp_min(lo[2]), warp_max(hi[0]), warp_max(hi[1]), warp_max(hi[2]), warp_max(hmax)};
    if (lane == 0)
        for (int k = 0; k < 7; k++) s[k][w] = v[k];
    __syncthreads();
    if (threadIdx.x < 7) {
        int k = threadIdx.x;
        float r = s[k][0];
        for (int j = 1; j < (int)(blockDim.x >> 5); j++) r = (k < 3) ? fminf(r, s[k][j]) : fmaxf(r, s[k][j]);
        if (k < 3) atomicMin(&bounds[k], f2ord(r));
        else atomicMax(&bounds[k], f2ord(r));
    }
}

// One thread: bounds -> grid parameters (mirror of orc_grid_params), then re-arm the bounds accumulator.
__global__ void k_grid_setup(uint32_t* bounds, sph_GridParams* g, int max_bits, int n) {
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) { lo[k] = ord2f(bounds[k]); hi[k] = ord2f(bounds[3 + k]); }
    float hmax = ord2f(bounds[6]);
    unsigned long long bitsum = *(unsigned long long*)(bounds + 8);
    float href = __uint_as_float(n > 0 ? (uint32_t)(bitsum / (unsigned long long)n) : 0u);
    float ext = fmaxf(fmaxf(__fsub_rn(hi[0], lo[0]), __fsub_rn(hi[1], lo[1])), __fsub_rn(hi[2], lo[2]));
    const int kStencilMax = 4;
    float reach = __fmul_rn(hmax, 2.002f);
    float cell = __fmul_rn(href, 2.002f);
    if (!(__fmul_rn(cell, (float)kStencilMax) >= reach)) cell = __fdiv_rn(reach, (float)kStencilMax);
    int bits = 0;
    for (; bits < max_bits; bits++)
        if (__fmul_rn(cell, (float)(1 << bits)) > ext) break;
    if (!(__fmul_rn(cell, (float)(1 << bits)) > ext)) cell = __fdiv_rn(__fmul_rn(ext, 1.0001f), (float)(1 << bits));
    if (!(cell > 0.0f)) cell = 1.0f;
    int S = 1;
    while (S < kStencilMax && !(__fmul_rn(cell, (float)S) >= reach)) S++;
    g->min[0] = lo[0]; g->min[1] = lo[1]; g->min[2] = lo[2];
    g->cell = cell; g->bits = bits; g->hmax = hmax; g->ext = ext; g->href = href; g->stencil = S;
    g->fine_scale = __fdiv_rn(1024.0f, __fmul_rn(cell, (float)(1 << bits)));
    bounds[0] = bounds[1] = bounds[2] = 0xffffffffu;
    bounds[3] = bounds[4] = bounds[5] = bounds[6] = 0u;
    bounds[8] = bounds[9] = 0u;
}

__global__ void __launch_bounds__(256) k_keys(const float4* __restrict__ posh, const sph_GridParams* __restrict__ g, int n,
                                              uint32_t* __restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = posh[i];
    float fs = g->fine_scale;
    int qx = (int)__fmul_rn(__fsub_rn(p.x, g->min[0]), fs);
    int qy = (int)__fmul_rn(__fsub_rn(p.y, g->min[1]), fs);
    int qz = (int)__fmul_rn(__fsub_rn(p.z, g->min[2]), fs);
    qx = min(max(qx, 0), 1023); qy = min(max(qy, 0), 1023); qz = min(max(qz, 0), 1023);
    keys[i] = expand10((uint32_t)qx) | (expand10((uint32_t)qy) << 1) | (expand10((uint32_t)qz) << 2);
}

// Gather the resident SoA into sorted order and emit the cell table from the sorted keys.
__global__ void __launch_bounds__(256) k_permute_cells(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                                                       const float4* __restrict__ posh_in, const float4* __restrict__ velm_in,
                                                       const uint32_t* __restrict__ orig_in, float4* __restrict__ posh_out,
                                                       float4* __restrict__ velm_out, uint32_t* __restrict__ orig_out,
                                                       float4* __restrict__ posm, float4* __restrict__ posc,
                                                       const sph_GridParams* __restrict__ g, int n,
                                                       uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_end,
                                                       uint32_t* __restrict__ cell_hmax) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t src = idx[i];
    float4 p = posh_in[src];
    float4 v = velm_in[src];
    posh_out[i] = p;
    velm_out[i] = v;
    orig_out[i] = orig_in[src];
    posm[i] = make_float4(p.x, p.y, p.z, v.w);
    posc[i] = make_float4(p.x, p.y, p.z, sph_keep_threshold(p.w));
    int shift = 3 * (10 - g->bits);
    uint32_t ck = keys[i] >> shift;
    if (i == 0 || (keys[i - 1] >> shift) != ck) cell_start[ck] = (uint32_t)i;
    if (i == n - 1 || (keys[i + 1] >> shift) != ck) cell_end[ck] = (uint32_t)(i + 1);
    atomicMax(&cell_hmax[ck], __float_as_uint(p.w));   // h > 0: the bit pattern orders like the value
}

}  // namespace

// smoothing-length update + bounds accumulation over `n` resident particles starting at posh / nown
int sph_launch_bounds_range(sphb200_ctx* c, float4* posh, const int32_t* nown, int n, bool update_h) {
    if (n <= 0) return SPH_OK;
    int blocks = min(sph_div_up(n, 256), c->sm_count * 8);
    k_smoothing_bounds<<<blocks, 256, 0, c->stream>>>(posh, nown, c->rr_table, n, update_h ? 1 : 0, c->bounds);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}
// bounds (of all n_total particles: a group all-reduces them first) -> grid parameters; re-arms the accumulator
int sph_launch_grid_setup(sphb200_ctx* c, int64_t n_total) {
    k_grid_setup<<<1, 1, 0, c->stream>>>(c->bounds, c->grid_d, c->grid_bits_max, (int)n_total);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}
int sph_launch_smoothing_bounds(sphb200_ctx* c, bool update_h) {
    int rc = sph_launch_bounds_range(c, c->posh[c->cur], c->nown, (int)c->n, update_h);
    if (rc) return rc;
    return sph_launch_grid_setup(c, c->n);
}
int sph_launch_keys(sphb200_ctx* c, const float4* posh, int n, uint32_t* keys) {
    if (n <= 0) return SPH_OK;
    k_keys<<<sph_div_up(n, 256), 256, 0, c->stream>>>(posh, c->grid_d, n, keys);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

int sph_launch_sort_and_cells(sphb200_ctx* c) {
    int n = (int)c->n;
    int in = c->cur, out = c->cur ^ 1;
    k_keys<<<sph_div_up(n, 256), 256, 0, c->stream>>>(c->posh[in], c->grid_d, n, c->keys[1]);
    SPH_LAUNCH_CHECK(c);
    { int rc = sph_launch_radix_sort(c, n, c->stream); if (rc) return rc; }   // (keys[1], iota) -> (keys[1], idx[1]), stable
    size_t ncell = (size_t)1 << (3 * c->grid_bits_max);   // cells of the largest grid this particle count can get
    SPH_CK(c, cudaMemsetAsync(c->cell_start, 0, ncell * sizeof(uint32_t), c->stream));
    SPH_CK(c, cudaMemsetAsync(c->cell_end, 0, ncell * sizeof(uint32_t), c->stream));
    SPH_CK(c, cudaMemsetAsync(c->cell_hmax, 0, ncell * sizeof(uint32_t), c->stream));
    k_permute_cells<<<sph_div_up(n, 256), 256, 0, c->stream>>>(c->keys[1], c->idx[1], c->posh[in], c->velm[in], c->orig[in],
                                                               c->posh[out], c->velm[out], c->orig[out], c->posm, c->posc, c->grid_d, n,
                                                               c->cell_start, c->cell_end, c->cell_hmax);
    SPH_LAUNCH_CHECK(c);
    c->cur = out;
    return SPH_OK;
}
