// kernels_sort.cu -- stable LSD radix sort of (30-bit Morton key, slot index) pairs and the stable multi-way partition
// that the Morton-range decomposition uses to route particles to their owner ranks.  Hand-written: no library kernel
// runs on the hot path.
//
// Replaces: KernelSystem's two single-threaded counting sorts by body A / body B (A/Systems/KernelSystem.cs:411-464,
// `SortPairsSTJob`; scheduled at :339-408) -- the reference sorts PAIRS so that every body finds its interactions; here
// the PARTICLES are sorted along the Morton curve once and every later stage (cell table, neighbor rows, LBVH) reads the
// sorted order.  The order is specified by oracle/sph_oracle.cpp (orc_sort_order: stable sort by key) and must match it
// bit for bit; a stable counting sort per 8-bit digit, least significant digit first, is exactly that.
//
// One pass = three launches:
//   k_digit_hist     every block counts the digits of its tile of SORT_TILE keys (shared-memory atomics) -> hist[d][block]
//   k_digit_rowscan  block d turns row d into exclusive prefixes over the tiles and writes total[d]
//   k_digit_scatter  every block ranks its tile stably -- warps own consecutive chunks, 32 consecutive keys per round,
//                    `match.any` groups equal digits, the group leader advances the warp's running count -- and writes
//                    (key, value) to base[d] + prefix[d][block] + rank.  base[] = exclusive scan of total[] (per block).
// Traffic per pass: 4 B (hist) + 8 B read + 8 B written per key; 4 passes cover 32 >= 30 key bits.
#include "ctx.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // keys per block
constexpr int SORT_WARPS = SORT_THREADS / 32;

// digit of element i: byte `shift/8` of its key, or -- partition mode -- the caller's bucket id
__device__ __forceinline__ int digit_of(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ bucket, int i, int shift) {
    return bucket ? (int)bucket[i] : (int)((keys[i] >> shift) & 0xffu);
}

__global__ void __launch_bounds__(SORT_THREADS) k_digit_hist(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ bucket,
                                                            int n, int shift, int nblocks, uint32_t* __restrict__ hist) {
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0u;
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int k = 0; k < SORT_ITEMS; k++) {
        const int i = base + k * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&cnt[digit_of(keys, bucket, i, shift)], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// exclusive scan of row d (one block per digit) in place; total[d] = row sum
__global__ void __launch_bounds__(256) k_digit_rowscan(uint32_t* __restrict__ hist, int nblocks, uint32_t* __restrict__ total) {
    __shared__ uint32_t wsum[8];
    uint32_t* row = hist + (size_t)blockIdx.x * nblocks;
    const int per = (nblocks + 255) / 256;
    const int b0 = threadIdx.x * per, b1 = min(b0 + per, nblocks);
    uint32_t s = 0u;
    for (int b = b0; b < b1; b++) s += row[b];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t off = 0u, all = 0u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < w) off += wsum[j];
        all += wsum[j];
    }
    uint32_t run = off + incl - s;
    for (int b = b0; b < b1; b++) {
        const uint32_t v = row[b];
        row[b] = run;
        run += v;
    }
    if (threadIdx.x == 0) total[blockIdx.x] = all;
}

// vals_in == nullptr: the value of element i is i (first pass / partition of an index range)
__global__ void __launch_bounds__(SORT_THREADS) k_digit_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                               const uint8_t* __restrict__ bucket, int n, int shift, int nblocks,
                                                               const uint32_t* __restrict__ hist, const uint32_t* __restrict__ total,
                                                               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t wcount[SORT_WARPS][256];
    __shared__ uint32_t dbase[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int k = threadIdx.x; k < SORT_WARPS * 256; k += SORT_THREADS) (&wcount[0][0])[k] = 0u;
    {   // base[d] = exclusive scan of the digit totals
        const uint32_t t = total[threadIdx.x];
        uint32_t incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        __shared__ uint32_t ws[8];
        if (lane == 31) ws[w] = incl;
        __syncthreads();
        uint32_t off = 0u;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j < w) off += ws[j];
        dbase[threadIdx.x] = off + incl - t + hist[(size_t)threadIdx.x * nblocks + blockIdx.x];
    }
    __syncthreads();
    // phase A: warp w owns keys [c0, c0 + 32*SORT_ITEMS) of the tile; rank of every key among the equal digits before it in
    // the warp's chunk
    const int c0 = blockIdx.x * SORT_TILE + w * (32 * SORT_ITEMS);
    int dig[SORT_ITEMS];
    uint32_t rnk[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const int i = c0 + r * 32 + lane;
        const bool live = i < n;
        const int d = live ? digit_of(keys_in, bucket, i, shift) : 256 + lane;   // idle lanes: unique pseudo-digits
        const unsigned peers = __match_any_sync(FULL, d);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0u;
        if (live && lane == leader) {
            pre = wcount[w][d];
            wcount[w][d] = pre + __popc(peers);
        }
        pre = __shfl_sync(FULL, pre, leader);
        dig[r] = live ? d : -1;
        rnk[r] = pre + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive scan over the warps, per digit, on top of the tile's global offset
        const int d = threadIdx.x;
        uint32_t run = dbase[d];
#pragma unroll
        for (int j = 0; j < SORT_WARPS; j++) {
            const uint32_t v = wcount[j][d];
            wcount[j][d] = run;
            run += v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        if (dig[r] < 0) continue;
        const int i = c0 + r * 32 + lane;
        const uint32_t pos = wcount[w][dig[r]] + rnk[r];
        SPH_DBG_IDX(pos, n);
        if (keys_out) keys_out[pos] = keys_in[i];
        vals_out[pos] = vals_in ? vals_in[i] : (uint32_t)i;
    }
}

}  // namespace

size_t sph_sort_hist_words(int64_t cap) { return (size_t)256 * (size_t)sph_div_up(cap, SORT_TILE) + 256; }

// exclusive scan of `rows` rows of `nblocks` counters each (row-major), totals[row] = row sum
int sph_launch_rowscan(sphb200_ctx* c, uint32_t* rows_d, int nblocks, int rows, uint32_t* totals, cudaStream_t stream) {
    k_digit_rowscan<<<rows, 256, 0, stream>>>(rows_d, nblocks, totals);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

// One stable counting pass.  bucket != nullptr: partition by bucket id (keys may be nullptr); total_out (256 words, device)
// receives the bucket sizes.
int sph_launch_digit_pass(sphb200_ctx* c, const uint32_t* keys_in, const uint32_t* vals_in, const uint8_t* bucket, int n, int shift,
                          uint32_t* keys_out, uint32_t* vals_out, uint32_t* total_out, cudaStream_t stream) {
    if (n <= 0) return SPH_OK;
    const int nblocks = sph_div_up(n, SORT_TILE);
    uint32_t* hist = c->sort_hist;
    uint32_t* total = total_out ? total_out : c->sort_hist + (size_t)256 * nblocks;
    k_digit_hist<<<nblocks, SORT_THREADS, 0, stream>>>(keys_in, bucket, n, shift, nblocks, hist);
    SPH_LAUNCH_CHECK(c);
    k_digit_rowscan<<<256, 256, 0, stream>>>(hist, nblocks, total);
    SPH_LAUNCH_CHECK(c);
    k_digit_scatter<<<nblocks, SORT_THREADS, 0, stream>>>(keys_in, vals_in, bucket, n, shift, nblocks, hist, total, keys_out, vals_out);
    SPH_LAUNCH_CHECK(c);
    return SPH_OK;
}

// Stable sort of (keys[1], iota) by the 30-bit key, in four 8-bit passes; the result is back in (keys[1], idx[1]).
int sph_launch_radix_sort(sphb200_ctx* c, int n, cudaStream_t stream) {
    int rc;
    if ((rc = sph_launch_digit_pass(c, c->keys[1], nullptr, nullptr, n, 0, c->keys[0], c->idx[0], nullptr, stream))) return rc;
    if ((rc = sph_launch_digit_pass(c, c->keys[0], c->idx[0], nullptr, n, 8, c->keys[1], c->idx[1], nullptr, stream))) return rc;
    if ((rc = sph_launch_digit_pass(c, c->keys[1], c->idx[1], nullptr, n, 16, c->keys[0], c->idx[0], nullptr, stream))) return rc;
    if ((rc = sph_launch_digit_pass(c, c->keys[0], c->idx[0], nullptr, n, 24, c->keys[1], c->idx[1], nullptr, stream))) return rc;
    return SPH_OK;
}
