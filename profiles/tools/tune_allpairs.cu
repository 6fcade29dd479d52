// tune_allpairs.cu -- standalone tuning harness for the all-pairs inner loop (not part of libsphb200.so).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tune_allpairs.cu -o tune_allpairs
// Prints ms and "20 flop/pair" TFLOP/s for a set of loop variants on N random particles; used to pick the shipped
// configuration of k_gravity_allpairs (results recorded in profiles/README.md).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cstdint>

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int TPT, int THREADS, int UNROLL, bool CAP, bool EQM, bool POT, int MINB, int ORDER = 0>
__global__ void __launch_bounds__(THREADS, MINB) k_ap(const float4* __restrict__ src, int n_src, int src_per_split,
                                                      const float4* __restrict__ posh, int nt, float4* __restrict__ part) {
    constexpr int TILE = THREADS > 256 ? THREADS : 256;
    __shared__ float4 tile[2][TILE];
    const int tid = threadIdx.x;
    const int tb = blockIdx.x * (THREADS * TPT);
    float xi[TPT], yi[TPT], zi[TPT], a2[TPT];
    float AX[TPT], AY[TPT], AZ[TPT], PH[TPT];
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * THREADS + tid;
        float4 p = posh[min(t, nt - 1)];
        xi[k] = p.x; yi[k] = p.y; zi[k] = p.z; a2[k] = p.w * p.w;
        AX[k] = AY[k] = AZ[k] = PH[k] = 0.f;
    }
    const int s0 = blockIdx.y * src_per_split;
    const int s1 = min(s0 + src_per_split, n_src);
    const int ntiles = (s1 - s0 + TILE - 1) / TILE;
    const float4 pad = make_float4(1.0e15f, 1.0e15f, 1.0e15f, 0.f);
    float4 nxt[TILE / THREADS];
#pragma unroll
    for (int q = 0; q < TILE / THREADS; q++) { int i = s0 + q * THREADS + tid; nxt[q] = i < s1 ? src[i] : pad; }
    for (int it = 0; it < ntiles; it++) {
        float4* buf = tile[it & 1];
#pragma unroll
        for (int q = 0; q < TILE / THREADS; q++) buf[q * THREADS + tid] = nxt[q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TILE / THREADS; q++) { int i = s0 + (it + 1) * TILE + q * THREADS + tid; nxt[q] = i < s1 ? src[i] : pad; }
        float ax[TPT], ay[TPT], az[TPT], ph[TPT];
#pragma unroll
        for (int k = 0; k < TPT; k++) ax[k] = ay[k] = az[k] = ph[k] = 0.f;
        const float4* tp = buf;
#pragma unroll 1
        for (int j = 0; j < TILE; j += UNROLL, tp += UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const float4 s = tp[u];
                if (ORDER == 1) {
                    float dx[TPT], dy[TPT], dz[TPT], r2[TPT], ri[TPT], g[TPT];
#pragma unroll
                    for (int k = 0; k < TPT; k++) dx[k] = xi[k] - s.x;
#pragma unroll
                    for (int k = 0; k < TPT; k++) dy[k] = yi[k] - s.y;
#pragma unroll
                    for (int k = 0; k < TPT; k++) dz[k] = zi[k] - s.z;
#pragma unroll
                    for (int k = 0; k < TPT; k++) r2[k] = dx[k] * dx[k];
#pragma unroll
                    for (int k = 0; k < TPT; k++) r2[k] = fmaf(dy[k], dy[k], r2[k]);
#pragma unroll
                    for (int k = 0; k < TPT; k++) r2[k] = fmaf(dz[k], dz[k], r2[k]);
#pragma unroll
                    for (int k = 0; k < TPT; k++) { if (CAP) r2[k] = fmaxf(r2[k], a2[k]); ri[k] = rsqrt_approx(r2[k]); }
#pragma unroll
                    for (int k = 0; k < TPT; k++) { if (EQM) { g[k] = ri[k] * ri[k]; } else { g[k] = ri[k] * ri[k]; ri[k] = s.w * ri[k]; } }
#pragma unroll
                    for (int k = 0; k < TPT; k++) { g[k] = g[k] * ri[k]; if (POT) ph[k] += ri[k]; }
#pragma unroll
                    for (int k = 0; k < TPT; k++) ax[k] = fmaf(dx[k], g[k], ax[k]);
#pragma unroll
                    for (int k = 0; k < TPT; k++) ay[k] = fmaf(dy[k], g[k], ay[k]);
#pragma unroll
                    for (int k = 0; k < TPT; k++) az[k] = fmaf(dz[k], g[k], az[k]);
                    continue;
                }
#pragma unroll
                for (int k = 0; k < TPT; k++) {
                    float dx = xi[k] - s.x, dy = yi[k] - s.y, dz = zi[k] - s.z;
                    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (CAP) r2 = fmaxf(r2, a2[k]);
                    float rinv = rsqrt_approx(r2);
                    float g;
                    if (EQM) {
                        g = rinv * rinv * rinv;
                        if (POT) ph[k] += rinv;
                    } else {
                        float mr = s.w * rinv;
                        g = mr * (rinv * rinv);
                        if (POT) ph[k] += mr;
                    }
                    ax[k] = fmaf(dx, g, ax[k]);
                    ay[k] = fmaf(dy, g, ay[k]);
                    az[k] = fmaf(dz, g, az[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < TPT; k++) { AX[k] += ax[k]; AY[k] += ay[k]; AZ[k] += az[k]; PH[k] += ph[k]; }
    }
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * THREADS + tid;
        if (t < nt) part[(size_t)blockIdx.y * nt + t] = make_float4(AX[k], AY[k], AZ[k], PH[k]);
    }
}

template <int TPT, int THREADS, int UNROLL, bool CAP, bool EQM, bool POT, int MINB, int ORDER = 0>
void run(const char* name, const float4* src, const float4* posh, float4* part, int n, int splits) {
    int tblocks = (n + THREADS * TPT - 1) / (THREADS * TPT);
    int per = ((n + splits - 1) / splits + 511) / 512 * 512;
    int sp = (n + per - 1) / per;
    dim3 grid(tblocks, sp);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_ap<TPT, THREADS, UNROLL, CAP, EQM, POT, MINB, ORDER><<<grid, THREADS>>>(src, n, per, posh, n, part);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_ap<TPT, THREADS, UNROLL, CAP, EQM, POT, MINB, ORDER>, THREADS, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_ap<TPT, THREADS, UNROLL, CAP, EQM, POT, MINB, ORDER>);
    printf("%-44s %8.3f ms  %6.2f TF(20/pair)  regs %3d  blocks/SM %d  grid %dx%d %s\n", name, best,
           20.0 * (double)n * n / (best * 1e-3) / 1e12, fa.numRegs, nb, tblocks, sp, e == cudaSuccess ? "" : cudaGetErrorString(e));
}


// ---- TMA (cp.async.bulk) variant: one elected thread streams 4 KB source tiles into a 4-stage shared-memory ring;
// full/empty mbarriers replace the block barrier.  Sources must be padded to a multiple of 256.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int TPT, int STAGES, bool EQM>
__global__ void __launch_bounds__(256) k_ap_tma(const float4* __restrict__ src, int n_src_padded, int src_per_split,
                                                const float4* __restrict__ posh, int nt, float4* __restrict__ part) {
    constexpr int TILE = 256, THREADS = 256;
    __shared__ __align__(128) float4 tile[STAGES][TILE];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    const int tid = threadIdx.x;
    const int tb = blockIdx.x * (THREADS * TPT);
    float xi[TPT], yi[TPT], zi[TPT];
    float AX[TPT], AY[TPT], AZ[TPT], PH[TPT];
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * THREADS + tid;
        float4 p = posh[min(t, nt - 1)];
        xi[k] = p.x; yi[k] = p.y; zi[k] = p.z;
        AX[k] = AY[k] = AZ[k] = PH[k] = 0.f;
    }
    const int s0 = blockIdx.y * src_per_split;
    const int s1 = min(s0 + src_per_split, n_src_padded);
    const int ntiles = (s1 - s0) / TILE;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < STAGES && s < ntiles; s++) {
            mbar_expect_tx(&full[s], TILE * 16);
            tma_load_1d(tile[s], src + s0 + s * TILE, TILE * 16, &full[s]);
        }
    for (int it = 0; it < ntiles; it++) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        float ax[TPT], ay[TPT], az[TPT], pp[TPT];
#pragma unroll
        for (int k = 0; k < TPT; k++) ax[k] = ay[k] = az[k] = pp[k] = 0.f;
        const float4* tp = tile[s];
#pragma unroll 1
        for (int j = 0; j < TILE; j += 8, tp += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float4 q = tp[u];
#pragma unroll
                for (int k = 0; k < TPT; k++) {
                    float dx = xi[k] - q.x, dy = yi[k] - q.y, dz = zi[k] - q.z;
                    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    float rinv = rsqrt_approx(r2);
                    float g;
                    if (EQM) { g = rinv * rinv * rinv; pp[k] += rinv; }
                    else { float mr = q.w * rinv; g = mr * (rinv * rinv); pp[k] += mr; }
                    ax[k] = fmaf(dx, g, ax[k]); ay[k] = fmaf(dy, g, ay[k]); az[k] = fmaf(dz, g, az[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < TPT; k++) { AX[k] += ax[k]; AY[k] += ay[k]; AZ[k] += az[k]; PH[k] += pp[k]; }
        mbar_arrive(&empty[s]);                       // this thread is done reading stage s
        if (tid == 0 && it + STAGES < ntiles) {       // refill it once everybody is
            mbar_wait(&empty[s], ph);
            mbar_expect_tx(&full[s], TILE * 16);
            tma_load_1d(tile[s], src + s0 + (it + STAGES) * TILE, TILE * 16, &full[s]);
        }
    }
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * THREADS + tid;
        if (t < nt) part[(size_t)blockIdx.y * nt + t] = make_float4(AX[k], AY[k], AZ[k], PH[k]);
    }
}

template <int TPT, int STAGES, bool EQM>
void run_tma(const char* name, const float4* src, const float4* posh, float4* part, int n, int splits) {
    int tblocks = (n + 256 * TPT - 1) / (256 * TPT);
    int per = ((n + splits - 1) / splits + 255) / 256 * 256;
    int sp = (n + per - 1) / per;
    dim3 grid(tblocks, sp);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_ap_tma<TPT, STAGES, EQM><<<grid, 256>>>(src, n, per, posh, n, part);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_ap_tma<TPT, STAGES, EQM>);
    printf("%-44s %8.3f ms  %6.2f TF(20/pair)  regs %3d  grid %dx%d %s\n", name, best, 20.0 * (double)n * n / (best * 1e-3) / 1e12,
           fa.numRegs, tblocks, sp, e == cudaSuccess ? "" : cudaGetErrorString(e));
}


// ---- packed-FP32 variant (Blackwell FFMA2/FADD2/FMUL2, PTX *.f32x2): one packed instruction evaluates the SAME target
// against TWO sources.  The tile is stored as source pairs, SoA inside the pair: A = (x0,x1,y0,y1), B = (z0,z1,m0,m1), so
// one LDS.128 pair feeds 2 sources x TPT targets; the target coordinate is the broadcast scalar operand (R.F32).
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int TPT, int THREADS, int UNROLL, bool CAP, bool EQM, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_ap2(const float4* __restrict__ src, int n_src, int src_per_split,
                                                       const float4* __restrict__ posh, int nt, float4* __restrict__ part) {
    constexpr int TILE = 256;
    static_assert(THREADS == 256, "one source per thread per tile");
    __shared__ __align__(16) float tile[2][TILE * 4];
    const int tid = threadIdx.x;
    const int tb = blockIdx.x * (THREADS * TPT);
    u64 xi[TPT], yi[TPT], zi[TPT];
    float a2[TPT];
    float AX[TPT], AY[TPT], AZ[TPT], PH[TPT];
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * THREADS + tid;
        float4 p = posh[min(t, nt - 1)];
        xi[k] = pk(p.x, p.x); yi[k] = pk(p.y, p.y); zi[k] = pk(p.z, p.z); a2[k] = p.w * p.w;
        AX[k] = AY[k] = AZ[k] = PH[k] = 0.f;
    }
    const int s0 = blockIdx.y * src_per_split;
    const int s1 = min(s0 + src_per_split, n_src);
    const int ntiles = (s1 - s0 + TILE - 1) / TILE;
    const float4 pad = make_float4(1.0e15f, 1.0e15f, 1.0e15f, 0.f);
    float4 nxt = (s0 + tid < s1) ? src[s0 + tid] : pad;
    const int so = (tid >> 1) * 8 + (tid & 1);
    for (int it = 0; it < ntiles; it++) {
        float* buf = tile[it & 1];
        buf[so] = nxt.x; buf[so + 2] = nxt.y; buf[so + 4] = nxt.z; buf[so + 6] = nxt.w;
        __syncthreads();
        { int i = s0 + (it + 1) * TILE + tid; nxt = i < s1 ? src[i] : pad; }
        u64 ax[TPT], ay[TPT], az[TPT], ph[TPT];
#pragma unroll
        for (int k = 0; k < TPT; k++) ax[k] = ay[k] = az[k] = ph[k] = pk(0.f, 0.f);
        const ulonglong2* tp = reinterpret_cast<const ulonglong2*>(buf);
#pragma unroll 1
        for (int j = 0; j < TILE / 2; j += UNROLL, tp += 2 * UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const ulonglong2 A = tp[2 * u], B = tp[2 * u + 1];   // A.x = (x0,x1) A.y = (y0,y1) B.x = (z0,z1) B.y = (m0,m1)
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
        for (int k = 0; k < TPT; k++) {
            float a, b;
            upk(ax[k], a, b); AX[k] += a + b;
            upk(ay[k], a, b); AY[k] += a + b;
            upk(az[k], a, b); AZ[k] += a + b;
            upk(ph[k], a, b); PH[k] += a + b;
        }
    }
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * THREADS + tid;
        if (t < nt) part[(size_t)blockIdx.y * nt + t] = make_float4(AX[k], AY[k], AZ[k], PH[k]);
    }
}

template <int TPT, int THREADS, int UNROLL, bool CAP, bool EQM, int MINB>
void run2(const char* name, const float4* src, const float4* posh, float4* part, int n, int splits) {
    int tblocks = (n + THREADS * TPT - 1) / (THREADS * TPT);
    int per = ((n + splits - 1) / splits + 511) / 512 * 512;
    int sp = (n + per - 1) / per;
    dim3 grid(tblocks, sp);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_ap2<TPT, THREADS, UNROLL, CAP, EQM, MINB><<<grid, THREADS>>>(src, n, per, posh, n, part);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_ap2<TPT, THREADS, UNROLL, CAP, EQM, MINB>, THREADS, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_ap2<TPT, THREADS, UNROLL, CAP, EQM, MINB>);
    printf("%-44s %8.3f ms  %6.2f TF(20/pair)  regs %3d  blocks/SM %d  grid %dx%d %s\n", name, best,
           20.0 * (double)n * n / (best * 1e-3) / 1e12, fa.numRegs, nb, tblocks, sp, e == cudaSuccess ? "" : cudaGetErrorString(e));
}


// ---- expansion variant (far tiles only): coordinates relative to the target block's centre c, sources carry
// s_j = |x_j - c|^2 (the mass slot is free for equal masses): r^2 = s_j + (s_i - 2 x_i'.x_j') = 3 FFMA2 + 1 FADD2 and the
// acceleration is accumulated as sum(g x_j') - x_i' sum(g): 11 FMA-pipe lane-ops per pair instead of 12.
template <int TPT, int UNROLL>
__global__ void __launch_bounds__(256, 1) k_ap2x(const float4* __restrict__ src, int n_src, int src_per_split,
                                                 const float4* __restrict__ posh, int nt, float4* __restrict__ part) {
    constexpr int TILE = 256, THREADS = 256;
    __shared__ __align__(16) float tile[2][TILE * 4];
    __shared__ float cbox[6][THREADS / 32];
    const int tid = threadIdx.x;
    const int tb = blockIdx.x * (THREADS * TPT);
    float px[TPT], py[TPT], pz[TPT];
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * THREADS + tid;
        float4 p = posh[min(t, nt - 1)];
        px[k] = p.x; py[k] = p.y; pz[k] = p.z;
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if ((tid & 31) == 0) for (int k = 0; k < 3; k++) { cbox[k][tid >> 5] = lo[k]; cbox[3 + k][tid >> 5] = hi[k]; }
    __syncthreads();
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float a = cbox[k][0], b = cbox[3 + k][0];
        for (int j = 1; j < THREADS / 32; j++) { a = fminf(a, cbox[k][j]); b = fmaxf(b, cbox[3 + k][j]); }
        c[k] = 0.5f * (a + b);
    }
    u64 mx[TPT], my[TPT], mz[TPT], si[TPT];   // -2 x_i' (packed twice), s_i
    float qx[TPT], qy[TPT], qz[TPT];
    float AX[TPT], AY[TPT], AZ[TPT], PH[TPT];
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        qx[k] = px[k] - c[0]; qy[k] = py[k] - c[1]; qz[k] = pz[k] - c[2];
        mx[k] = pk(-2.f * qx[k], -2.f * qx[k]); my[k] = pk(-2.f * qy[k], -2.f * qy[k]); mz[k] = pk(-2.f * qz[k], -2.f * qz[k]);
        float s = qx[k] * qx[k] + qy[k] * qy[k] + qz[k] * qz[k];
        si[k] = pk(s, s);
        AX[k] = AY[k] = AZ[k] = PH[k] = 0.f;
    }
    const int s0 = blockIdx.y * src_per_split;
    const int s1 = min(s0 + src_per_split, n_src);
    const int ntiles = (s1 - s0 + TILE - 1) / TILE;
    const float4 pad = make_float4(1.0e15f, 1.0e15f, 1.0e15f, 0.f);
    float4 nxt = (s0 + tid < s1) ? src[s0 + tid] : pad;
    const int so = (tid >> 1) * 8 + (tid & 1);
    for (int it = 0; it < ntiles; it++) {
        float* buf = tile[it & 1];
        {
            float x = nxt.x - c[0], y = nxt.y - c[1], z = nxt.z - c[2];
            buf[so] = x; buf[so + 2] = y; buf[so + 4] = z; buf[so + 6] = x * x + y * y + z * z;
        }
        __syncthreads();
        { int i = s0 + (it + 1) * TILE + tid; nxt = i < s1 ? src[i] : pad; }
        u64 ax[TPT], ay[TPT], az[TPT], sg[TPT], ph[TPT];
#pragma unroll
        for (int k = 0; k < TPT; k++) ax[k] = ay[k] = az[k] = sg[k] = ph[k] = pk(0.f, 0.f);
        const ulonglong2* tp = reinterpret_cast<const ulonglong2*>(buf);
#pragma unroll 1
        for (int j = 0; j < TILE / 2; j += UNROLL, tp += 2 * UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const ulonglong2 A = tp[2 * u], B = tp[2 * u + 1];   // A = (x0,x1),(y0,y1)  B = (z0,z1),(s0,s1)
#pragma unroll
                for (int k = 0; k < TPT; k++) {
                    u64 r2 = add2(fma2(mx[k], A.x, fma2(my[k], A.y, fma2(mz[k], B.x, B.y))), si[k]);
                    float r2a, r2b;
                    upk(r2, r2a, r2b);
                    u64 rinv = pk(rsqrt_approx(r2a), rsqrt_approx(r2b));
                    u64 g = mul2(mul2(rinv, rinv), rinv);
                    ph[k] = add2(ph[k], rinv);
                    sg[k] = add2(sg[k], g);
                    ax[k] = fma2(A.x, g, ax[k]);
                    ay[k] = fma2(A.y, g, ay[k]);
                    az[k] = fma2(B.x, g, az[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < TPT; k++) {   // d = x_i - x_j = x_i' - x_j': sum(g d) = x_i' sum(g) - sum(g x_j')
            float a, b, ga, gb;
            upk(sg[k], ga, gb); const float gs = ga + gb;
            upk(ax[k], a, b); AX[k] += fmaf(qx[k], gs, -(a + b));
            upk(ay[k], a, b); AY[k] += fmaf(qy[k], gs, -(a + b));
            upk(az[k], a, b); AZ[k] += fmaf(qz[k], gs, -(a + b));
            upk(ph[k], a, b); PH[k] += a + b;
        }
    }
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        int t = tb + k * THREADS + tid;
        if (t < nt) part[(size_t)blockIdx.y * nt + t] = make_float4(AX[k], AY[k], AZ[k], PH[k]);
    }
}

template <int TPT, int UNROLL>
void run2x(const char* name, const float4* src, const float4* posh, float4* part, int n, int splits) {
    int tblocks = (n + 256 * TPT - 1) / (256 * TPT);
    int per = ((n + splits - 1) / splits + 511) / 512 * 512;
    int sp = (n + per - 1) / per;
    dim3 grid(tblocks, sp);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_ap2x<TPT, UNROLL><<<grid, 256>>>(src, n, per, posh, n, part);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_ap2x<TPT, UNROLL>);
    printf("%-44s %8.3f ms  %6.2f TF(20/pair)  regs %3d  grid %dx%d %s\n", name, best, 20.0 * (double)n * n / (best * 1e-3) / 1e12,
           fa.numRegs, tblocks, sp, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// FMA-pipe peak: scalar FFMA chains vs packed FFMA2 chains (both counted as 2 flop per fp32 lane-op).
template <bool PACKED>
__global__ void __launch_bounds__(256) k_peak(float* out, int iters, float a, float b) {
    if (PACKED) {
        u64 v[8], a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = pk((float)(threadIdx.x + k), (float)k);
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = fma2(v[k], a2, b2);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) { float x, y; upk(v[k], x, y); s += x + y; }
        if (s == 12345.678f) out[0] = s;
    } else {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (float)(threadIdx.x + k);
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = fmaf(v[k], a, b);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) s += v[k];
        if (s == 12345.678f) out[0] = s;
    }
}
template <bool PACKED>
void run_peak(const char* name, float* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * 8, iters = 4096;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_peak<PACKED><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 64.0 * (PACKED ? 2.0 : 1.0) * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    printf("%-44s %6.2f TFLOP/s\n", name, best);
}

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 262144;
    std::vector<float4> h(n), p(n);
    srand(1);
    for (int i = 0; i < n; i++) {
        float x = rand() / (float)RAND_MAX * 100, y = rand() / (float)RAND_MAX * 100, z = rand() / (float)RAND_MAX * 100;
        h[i] = make_float4(x, y, z, 1.0f / n);
        p[i] = make_float4(x, y, z, 0.5f);
    }
    float4 *src, *posh, *part;
    cudaMalloc(&src, n * 16); cudaMalloc(&posh, n * 16); cudaMalloc(&part, (size_t)n * 16 * 64);
    cudaMemcpy(src, h.data(), n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(posh, p.data(), n * 16, cudaMemcpyHostToDevice);
    //            TPT THR UNR  CAP    EQM    POT   MINB
    run<4, 256, 8, true, false, true, 1>("shipped: tpt4 t256 u8 cap mass pot", src, posh, part, n, 28);
    run<4, 256, 8, false, false, true, 1>("no cap (far tiles)", src, posh, part, n, 28);
    run<4, 256, 8, true, true, true, 1>("equal mass", src, posh, part, n, 28);
    run<4, 256, 8, false, true, true, 1>("equal mass, no cap", src, posh, part, n, 28);
    run<4, 256, 8, false, true, false, 1>("equal mass, no cap, no potential", src, posh, part, n, 28);
    run<2, 256, 8, true, false, true, 1>("tpt2 t256", src, posh, part, n, 14);
    run<8, 128, 8, true, false, true, 1>("tpt8 t128", src, posh, part, n, 28);
    run<8, 256, 4, true, false, true, 1>("tpt8 t256 u4", src, posh, part, n, 56);
    run<4, 128, 8, true, false, true, 1>("tpt4 t128", src, posh, part, n, 14);
    run<4, 256, 4, true, false, true, 1>("tpt4 t256 u4", src, posh, part, n, 28);
    run<4, 256, 16, true, false, true, 1>("tpt4 t256 u16", src, posh, part, n, 28);
    run<4, 256, 8, true, false, true, 2>("tpt4 t256 minb2", src, posh, part, n, 28);
    run<4, 256, 8, true, false, true, 4>("tpt4 t256 minb4 (<=64 regs)", src, posh, part, n, 28);
    run<4, 512, 8, true, false, true, 1>("tpt4 t512", src, posh, part, n, 56);
    run<6, 128, 8, true, false, true, 1>("tpt6 t128", src, posh, part, n, 21);
    run<3, 256, 8, true, false, true, 1>("tpt3 t256", src, posh, part, n, 21);
    run<4, 256, 8, false, true, true, 1, 1>("equal mass, no cap, component-major", src, posh, part, n, 28);
    run<8, 128, 4, false, true, true, 1, 1>("eqm nocap comp-major tpt8 t128 u4", src, posh, part, n, 28);
    run<2, 256, 8, false, true, true, 1, 1>("eqm nocap comp-major tpt2", src, posh, part, n, 14);
    run<4, 128, 8, false, true, true, 1>("eqm nocap tpt4 t128", src, posh, part, n, 14);
    run<2, 256, 8, false, true, true, 1>("eqm nocap tpt2 t256", src, posh, part, n, 14);
    run<2, 128, 16, false, true, true, 1>("eqm nocap tpt2 t128 u16", src, posh, part, n, 7);
    run<8, 128, 8, false, true, true, 1>("eqm nocap tpt8 t128", src, posh, part, n, 28);
    run_peak<false>("FMA peak: scalar FFMA chains", (float*)part);
    run_peak<true>("FMA peak: packed FFMA2 chains", (float*)part);
    //    TPT THR UNR CAP    EQM   MINB
    run2<4, 256, 4, false, true, 1>("f32x2: eqm nocap tpt4 u4", src, posh, part, n, 28);
    run2<4, 256, 2, false, true, 1>("f32x2: eqm nocap tpt4 u2", src, posh, part, n, 28);
    run2<4, 256, 8, false, true, 1>("f32x2: eqm nocap tpt4 u8", src, posh, part, n, 28);
    run2<2, 256, 4, false, true, 1>("f32x2: eqm nocap tpt2 u4", src, posh, part, n, 14);
    run2<2, 256, 8, false, true, 1>("f32x2: eqm nocap tpt2 u8", src, posh, part, n, 14);
    run2<3, 256, 4, false, true, 1>("f32x2: eqm nocap tpt3 u4", src, posh, part, n, 21);
    run2<6, 256, 2, false, true, 1>("f32x2: eqm nocap tpt6 u2", src, posh, part, n, 42);
    run2<8, 256, 2, false, true, 1>("f32x2: eqm nocap tpt8 u2", src, posh, part, n, 56);
    run2<4, 256, 4, true, true, 1>("f32x2: eqm cap tpt4 u4", src, posh, part, n, 28);
    run2<4, 256, 4, false, false, 1>("f32x2: mass nocap tpt4 u4", src, posh, part, n, 28);
    run2<4, 256, 4, true, false, 1>("f32x2: mass cap tpt4 u4", src, posh, part, n, 28);
    {
        std::vector<float4> q1(n), q2(n);
        k_ap<4, 256, 8, false, false, true, 1><<<dim3((n + 1023) / 1024, 1), 256>>>(src, n, n, posh, n, part);
        cudaMemcpy(q1.data(), part, n * 16, cudaMemcpyDeviceToHost);
        k_ap2<4, 256, 4, false, false, 1><<<dim3((n + 1023) / 1024, 1), 256>>>(src, n, n, posh, n, part);
        cudaMemcpy(q2.data(), part, n * 16, cudaMemcpyDeviceToHost);
        double md = 0, mw = 0;
        for (int i = 0; i < n; i++) {
            md = fmax(md, fabs(q1[i].x - q2[i].x) / (fabs(q1[i].x) + 1e-30));
            mw = fmax(mw, fabs(q1[i].w - q2[i].w) / (fabs(q1[i].w) + 1e-30));
        }
        printf("max rel diff plain vs f32x2: x %.3e  phi %.3e  %s\n", md, mw, cudaGetErrorString(cudaGetLastError()));
    }
    run2x<4, 4>("f32x2 expansion (far tiles): eqm tpt4 u4", src, posh, part, n, 28);
    run2x<4, 2>("f32x2 expansion (far tiles): eqm tpt4 u2", src, posh, part, n, 28);
    run2x<2, 4>("f32x2 expansion (far tiles): eqm tpt2 u4", src, posh, part, n, 14);
    run2x<3, 4>("f32x2 expansion (far tiles): eqm tpt3 u4", src, posh, part, n, 21);
    run_tma<4, 4, true>("TMA ring: eqm nocap tpt4 4 stages", src, posh, part, n, 28);
    run_tma<4, 2, true>("TMA ring: eqm nocap tpt4 2 stages", src, posh, part, n, 28);
    run_tma<4, 4, false>("TMA ring: mass nocap tpt4 4 stages", src, posh, part, n, 28);
    // checksum of the last variants vs the plain kernel
    std::vector<float4> r1(n), r2(n);
    k_ap<4, 256, 8, false, false, true, 1><<<dim3((n + 1023) / 1024, 1), 256>>>(src, n, n, posh, n, part);
    cudaMemcpy(r1.data(), part, n * 16, cudaMemcpyDeviceToHost);
    k_ap_tma<4, 4, false><<<dim3((n + 1023) / 1024, 1), 256>>>(src, n, n, posh, n, part);
    cudaMemcpy(r2.data(), part, n * 16, cudaMemcpyDeviceToHost);
    double md = 0; for (int i = 0; i < n; i++) md = fmax(md, fabs(r1[i].x - r2[i].x) / (fabs(r1[i].x) + 1e-30));
    printf("max rel diff plain vs TMA (x component): %.3e  %s\n", md, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
