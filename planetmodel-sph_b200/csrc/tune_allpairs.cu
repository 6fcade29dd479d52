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
