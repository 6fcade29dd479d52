// kernels_group.cu -- device side of the Morton-range domain decomposition (group.cu drives it).
//
// Replaces: nothing the reference has -- its only parallelism is shared-memory job threads (UP/Collision/World/
// Broadphase.cs:163) and it is capped at 2^24-2 bodies (UP/Dynamics/Simulation/Scheduler.cs:24-41).  The decomposition keeps
// the single-GPU semantics exactly: the global sorted order is the stable sort of the previous global order by Morton key
// (oracle: orc_sort_order), every rank owns a contiguous key range of it, and each kernel of the hot path sees, for its own
// targets, the same sources in the same order as on one GPU -- so results are bit-identical to the single-GPU step.
//
//   k_bin_hist / k_splitters / k_dest   ownership: histogram of the new keys over 2^18 key-prefix bins (all-reduced), splitters
//                                       at bin boundaries balancing the *work* of the last step, destination rank per particle
//   k_mig_pack / k_mig_keys             migration records (48 B) in stable per-destination order; keys of the received set
//   k_halo_mask / count / scatter / pack  halo: particles whose cell stencil touches a cell owned by another rank, compacted
//                                       stably per destination (32 B records: posh + m, body index, key)
//   k_assemble_ext                      resident "extended" set = [halo of lower ranks | own | halo of higher ranks] -- a sorted
//                                       subsequence of the global order -- with its SoA arrays and cell table
//   k_boundary                          first / last particles of the rank (buckets that straddle a rank boundary)
//   k_let_box / mask / pack / scatter   locally essential tree: only the walk nodes another rank can reach travel to it
//   k_result_* / k_body_dest            redistribution of the results to body-order slices for the download
//   k_reduce_ranks                      all-reduce of the in-process transport (single process, ranks on one or more devices)
#include "ctx.cuh"
#include "group.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// bins are key prefixes, never finer than the grid cells (a cell belongs to exactly one rank)
__device__ __forceinline__ int bin_bits(int grid_bits) { return min(SPH_BIN_BITS, 3 * grid_bits); }

// Weight of one particle in the balance: what its density/force walks and its gravity walk cost in the last step
// (C4 on B200: ~28 ps per neighbor, ~3.3 ps per tree interaction).
__device__ __forceinline__ uint32_t work_of(int ncount, int npart, int napprox) {
    return 8u * (uint32_t)(max(ncount, 0) + 1) + (uint32_t)max(npart, 0) + (uint32_t)max(napprox, 0);
}

__global__ void __launch_bounds__(256) k_bin_hist(const uint32_t* __restrict__ keys, const int32_t* __restrict__ ncount,
                                                  const int32_t* __restrict__ npart, const int32_t* __restrict__ napprox, int n,
                                                  const sph_GridParams* __restrict__ g, uint32_t* __restrict__ hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int shift = 30 - bin_bits(g->bits);
    const bool live = i < n;
    const uint32_t bin = live ? keys[i] >> shift : 0xffffff00u + lane;
    const uint32_t w = live ? work_of(ncount[i], npart[i], napprox[i]) : 0u;
    // the keys arrive nearly sorted: lanes of a warp mostly share a bin -- one atomic per distinct bin
    const unsigned peers = __match_any_sync(FULL, bin);
    const uint32_t wsum = __reduce_add_sync(peers, w);
    if (live && lane == __ffs(peers) - 1) {
        atomicAdd(&hist[bin], (uint32_t)__popc(peers));
        atomicAdd(&hist[SPH_NBINS + bin], wsum);
    }
}

// Splitter r (1..world-1) = the first bin b whose exclusive work prefix reaches r/world of the total work; rank r owns bins
// [sbin[r], sbin[r+1]) and global sorted slots [g0[r], g0[r+1]).  out: g0[0..world], sbin[0..world].
// Two levels: k_bin_partials sums runs of 1024 bins (coalesced, 256 blocks); k_splitters scans the 256 run sums, locates the
// run every splitter falls into and scans only that run.
constexpr int SPL_RUN = 1024, SPL_NRUN = SPH_NBINS / SPL_RUN;

__global__ void __launch_bounds__(SPL_RUN) k_bin_partials(const uint32_t* __restrict__ hist, unsigned long long* __restrict__ part) {
    __shared__ unsigned long long sc[32], sw[32];
    const int b = blockIdx.x * SPL_RUN + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long c = hist[b], wk = hist[SPH_NBINS + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(FULL, c, o); wk += __shfl_xor_sync(FULL, wk, o); }
    if (lane == 0) { sc[w] = c; sw[w] = wk; }
    __syncthreads();
    if (w == 0) {
        c = sc[lane]; wk = sw[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(FULL, c, o); wk += __shfl_xor_sync(FULL, wk, o); }
        if (lane == 0) { part[blockIdx.x] = c; part[SPL_NRUN + blockIdx.x] = wk; }
    }
}

__global__ void __launch_bounds__(SPL_RUN) k_splitters(const uint32_t* __restrict__ hist, const unsigned long long* __restrict__ part,
                                                       int world, int64_t* __restrict__ out) {
    __shared__ unsigned long long rw[SPL_NRUN + 1], rc[SPL_NRUN + 1];   // exclusive prefixes of the runs
    __shared__ unsigned long long wsum_w[32], wsum_c[32];
    __shared__ int run_of;
    const int tx = threadIdx.x, lane = tx & 31, wp = tx >> 5;
    if (tx == 0) {
        unsigned long long aw = 0ull, ac = 0ull;
        for (int k = 0; k < SPL_NRUN; k++) { rc[k] = ac; rw[k] = aw; ac += part[k]; aw += part[SPL_NRUN + k]; }
        rc[SPL_NRUN] = ac; rw[SPL_NRUN] = aw;
        for (int r = 0; r < world; r++) { out[r] = 0; out[world + 1 + r] = 0; }
        out[world] = (int64_t)ac; out[2 * world + 1] = SPH_NBINS;
    }
    __syncthreads();
    const unsigned long long tot = rw[SPL_NRUN];
    for (int r = 1; r < world; r++) {
        const unsigned long long target = tot / (unsigned long long)world * r + (tot % (unsigned long long)world) * r / world;
        if (target == 0ull) continue;                       // block-uniform
        if (tx < SPL_NRUN && rw[tx] < target && target <= rw[tx + 1]) run_of = tx;   // exactly one run
        __syncthreads();
        const int k = run_of, b = k * SPL_RUN + tx;
        const unsigned long long w = hist[SPH_NBINS + b], c = hist[b];
        unsigned long long iw = w, ic = c;                   // inclusive scan over the run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long vw = __shfl_up_sync(FULL, iw, o), vc = __shfl_up_sync(FULL, ic, o);
            if (lane >= o) { iw += vw; ic += vc; }
        }
        if (lane == 31) { wsum_w[wp] = iw; wsum_c[wp] = ic; }
        __syncthreads();
        unsigned long long bw = rw[k], bc = rc[k];
        for (int j = 0; j < wp; j++) { bw += wsum_w[j]; bc += wsum_c[j]; }
        const unsigned long long pw = bw + iw, before = pw - w;
        if (before < target && target <= pw) { out[r] = (int64_t)(bc + ic); out[world + 1 + r] = b + 1; }
        __syncthreads();
    }
}

// Owners found among the bins within `sb` bins of every bin (bit q: rank q): a particle whose bin sees only its own rank
// cannot be in any halo (k_halo_mask skips it).  sb = stencil cells rounded up to whole bins.
__global__ void __launch_bounds__(256) k_bin_owner_mask(const sph_GridParams* __restrict__ g, const int64_t* __restrict__ split, int world,
                                                        uint32_t* __restrict__ bmask) {
    __shared__ int sbin[SPH_MAX_RANKS + 1];
    if (threadIdx.x <= world) sbin[threadIdx.x] = (int)split[world + 1 + threadIdx.x];
    __syncthreads();
    const int bb = bin_bits(g->bits);                 // a multiple of 3: bins are cubes of 2^(bits - bb/3) cells
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (1 << bb)) return;
    const int per_axis = bb / 3, bdim = 1 << per_axis, cells = 1 << (g->bits - per_axis);
    const int sb = (g->stencil + cells - 1) / cells;
    const int bx = (int)compact10((uint32_t)b), by = (int)compact10((uint32_t)b >> 1), bz = (int)compact10((uint32_t)b >> 2);
    uint32_t m = 0u;
    for (int oz = -sb; oz <= sb; oz++)
        for (int oy = -sb; oy <= sb; oy++)
            for (int ox = -sb; ox <= sb; ox++) {
                const int nx = bx + ox, ny = by + oy, nz = bz + oz;
                if (nx < 0 || ny < 0 || nz < 0 || nx >= bdim || ny >= bdim || nz >= bdim) continue;
                const int nb = (int)(expand10((uint32_t)nx) | (expand10((uint32_t)ny) << 1) | (expand10((uint32_t)nz) << 2));
                int r = 0;
                for (int q = 1; q < world; q++) r += sbin[q] <= nb ? 1 : 0;
                m |= 1u << r;
            }
    bmask[b] = m;
}

__global__ void __launch_bounds__(256) k_dest(const uint32_t* __restrict__ keys, int n, const sph_GridParams* __restrict__ g,
                                              const int64_t* __restrict__ split, int world, uint8_t* __restrict__ dest) {
    __shared__ int sbin[SPH_MAX_RANKS + 1];
    if (threadIdx.x <= world) sbin[threadIdx.x] = (int)split[world + 1 + threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bin = (int)(keys[i] >> (30 - bin_bits(g->bits)));
    int r = 0;
    for (int q = 1; q < world; q++) r += sbin[q] <= bin ? 1 : 0;
    dest[i] = (uint8_t)r;
}

// migration record: [0] posh, [1] velm, [2] (body index, own-support count, key, -)
__global__ void __launch_bounds__(256) k_mig_pack(const float4* __restrict__ posh, const float4* __restrict__ velm,
                                                  const uint32_t* __restrict__ orig, const int32_t* __restrict__ nown,
                                                  const uint32_t* __restrict__ keys, const uint32_t* __restrict__ perm, int n,
                                                  uint4* __restrict__ rec) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t i = perm[k];
    const float4 p = posh[i], v = velm[i];
    rec[3 * (size_t)k] = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), __float_as_uint(p.w));
    rec[3 * (size_t)k + 1] = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
    rec[3 * (size_t)k + 2] = make_uint4(orig[i], (uint32_t)nown[i], keys[i], 0u);
}

// the own particles in sorted order, as soon as the local sort is done: gravity sources and keys straight into the rank's segment of
// the global arrays (what the all-gather publishes) and the (posh, velm) records the LBVH build reads
__global__ void __launch_bounds__(256) k_sorted_sources(const uint4* __restrict__ rec, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ keys,
                                                        int n, float4* __restrict__ posm, uint32_t* __restrict__ keys_out, float4* __restrict__ tposh,
                                                        float4* __restrict__ tvelm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t s = idx[i];
    const uint4 a = rec[3 * s], b = rec[3 * s + 1];
    const float4 p = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
    const float4 v = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
    posm[i] = make_float4(p.x, p.y, p.z, v.w);
    if (keys_out) keys_out[i] = keys[i];
    tposh[i] = p; tvelm[i] = v;
}

__global__ void __launch_bounds__(256) k_mig_keys(const uint4* __restrict__ rec, int n, uint32_t* __restrict__ keys) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) keys[k] = rec[3 * (size_t)k + 2].z;
}

// Halo destinations of every own particle (sorted order): bit q set <=> some cell within the stencil of the particle's
// cell belongs to rank q != me.  A pair (i,j) can only interact if their cells are within S cells of each other on every
// axis (S * cell >= 2.002 h_max, k_grid_setup), so this is a superset of what rank q's neighbor search can reach.
// The lanes of a warp that share a cell split its stencil among themselves.
__global__ void __launch_bounds__(256) k_halo_mask(const uint32_t* __restrict__ keys, int n, const sph_GridParams* __restrict__ g,
                                                   const int64_t* __restrict__ split, int world, int me, const uint32_t* __restrict__ bmask,
                                                   uint32_t* __restrict__ mask) {
    __shared__ int sbin[SPH_MAX_RANKS + 1];
    if (threadIdx.x <= world) sbin[threadIdx.x] = (int)split[world + 1 + threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int bits = g->bits, S = g->stencil, dim = 1 << bits, nst = 2 * S + 1, nst3 = nst * nst * nst;
    const int cshift = 3 * (10 - bits), bshift = 3 * bits - bin_bits(bits);
    const bool live = i < n;
    // interior particles (the bins around theirs all belong to this rank) are most of the domain: whole warps leave here
    const bool maybe = live && (bmask[keys[i] >> (30 - bin_bits(bits))] & ~(1u << me)) != 0u;
    if (!__any_sync(FULL, maybe)) { if (live) mask[i] = 0u; return; }
    const uint32_t ck = live ? keys[i] >> cshift : 0xffffff00u + lane;
    const unsigned peers = __match_any_sync(FULL, ck);
    const int np = __popc(peers), pr = __popc(peers & ((1u << lane) - 1u));
    uint32_t m = 0u;
    if (live) {
        const int cx = (int)compact10(ck), cy = (int)compact10(ck >> 1), cz = (int)compact10(ck >> 2);
        for (int k = pr; k < nst3; k += np) {
            const int oz = k / (nst * nst), rem = k - oz * nst * nst, oy = rem / nst, ox = rem - oy * nst;
            const int nx = cx + ox - S, ny = cy + oy - S, nz = cz + oz - S;
            if (nx < 0 || ny < 0 || nz < 0 || nx >= dim || ny >= dim || nz >= dim) continue;
            const uint32_t nk = expand10((uint32_t)nx) | (expand10((uint32_t)ny) << 1) | (expand10((uint32_t)nz) << 2);
            const int bin = (int)(nk >> bshift);
            int r = 0;
            for (int q = 1; q < world; q++) r += sbin[q] <= bin ? 1 : 0;
            m |= 1u << r;
        }
    }
    m = __reduce_or_sync(peers, m);
    if (live) mask[i] = m & ~(1u << me);
}

constexpr int HC_THREADS = 256;
constexpr int HC_ITEMS = 8;
constexpr int HC_TILE = HC_THREADS * HC_ITEMS;

// per tile and destination: number of particles to send (cnt[q][tile])
__global__ void __launch_bounds__(HC_THREADS) k_halo_count(const uint32_t* __restrict__ mask, int n, int world, int ntiles,
                                                           uint32_t* __restrict__ cnt) {
    __shared__ uint32_t c[SPH_MAX_RANKS];
    if (threadIdx.x < SPH_MAX_RANKS) c[threadIdx.x] = 0u;
    __syncthreads();
    const int base = blockIdx.x * HC_TILE;
    uint32_t m[HC_ITEMS];
#pragma unroll
    for (int k = 0; k < HC_ITEMS; k++) {
        const int i = base + k * HC_THREADS + threadIdx.x;
        m[k] = i < n ? mask[i] : 0u;
    }
    for (int q = 0; q < world; q++) {
        int s = 0;
#pragma unroll
        for (int k = 0; k < HC_ITEMS; k++) s += (m[k] >> q) & 1u;
        s = __reduce_add_sync(FULL, s);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&c[q], (uint32_t)s);
    }
    __syncthreads();
    if (threadIdx.x < world) cnt[(size_t)threadIdx.x * ntiles + blockIdx.x] = c[threadIdx.x];
}

// stable compaction: list of destination q = [qbase[q] + prefix[q][tile] + rank in tile), ascending particle index.
// A tile is walked in index order: round k holds indices base + k*256 + thread.
__global__ void __launch_bounds__(HC_THREADS) k_halo_scatter(const uint32_t* __restrict__ mask, int n, int world, int ntiles,
                                                             const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ total,
                                                             uint32_t* __restrict__ list) {
    __shared__ uint32_t run[SPH_MAX_RANKS];      // next free position per destination
    __shared__ uint32_t wcnt[HC_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < world) {
        uint32_t qb = 0u;
        for (int q = 0; q < (int)threadIdx.x; q++) qb += total[q];
        run[threadIdx.x] = qb + prefix[(size_t)threadIdx.x * ntiles + blockIdx.x];
    }
    __syncthreads();
    const int base = blockIdx.x * HC_TILE;
    for (int k = 0; k < HC_ITEMS; k++) {
        const int i = base + k * HC_THREADS + threadIdx.x;
        const uint32_t m = i < n ? mask[i] : 0u;
        for (int q = 0; q < world; q++) {
            const bool f = (m >> q) & 1u;
            const unsigned bal = __ballot_sync(FULL, f);
            if (lane == 0) wcnt[w] = __popc(bal);
            __syncthreads();
            uint32_t off = 0u, all = 0u;
#pragma unroll
            for (int j = 0; j < HC_THREADS / 32; j++) { if (j < w) off += wcnt[j]; all += wcnt[j]; }
            if (f) list[run[q] + off + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)i;
            __syncthreads();
            if (threadIdx.x == 0) run[q] += all;
        }
        __syncthreads();
    }
}

// halo record: [0] posh, [1] (m, body index, key, -); the particle of sorted rank i sits in rec[idx[i]]
__global__ void __launch_bounds__(256) k_halo_pack(const uint4* __restrict__ rec, const uint32_t* __restrict__ idx,
                                                   const uint32_t* __restrict__ keys, const uint32_t* __restrict__ list, int n,
                                                   uint4* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t i = list[k];
    const size_t s = idx[i];
    out[2 * (size_t)k] = rec[3 * s];
    out[2 * (size_t)k + 1] = make_uint4(rec[3 * s + 1].w, rec[3 * s + 2].x, keys[i], 0u);
}

// Extended resident set [low halo | own | high halo]: SoA arrays, gravity/neighbor records and the cell table.
__global__ void __launch_bounds__(256) k_assemble_ext(const uint4* __restrict__ rec, const uint32_t* __restrict__ idx,
                                                      const uint32_t* __restrict__ keys_own, const uint4* __restrict__ halo, int low, int nown_,
                                                      int next, const sph_GridParams* __restrict__ g, float4* __restrict__ posh,
                                                      float4* __restrict__ velm, uint32_t* __restrict__ orig, int32_t* __restrict__ nown,
                                                      uint32_t* __restrict__ keys_ext, float4* __restrict__ posm, float4* __restrict__ posc,
                                                      uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_end,
                                                      uint32_t* __restrict__ cell_hmax) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= next) return;
    auto key_of = [&](int x) -> uint32_t {
        if (x < low) return halo[2 * (size_t)x + 1].z;
        if (x < low + nown_) return keys_own[x - low];
        return halo[2 * (size_t)(x - nown_) + 1].z;
    };
    float4 p, v;
    uint32_t o, key;
    int no = 0;
    if (e >= low && e < low + nown_) {
        const size_t s = idx[e - low];
        const uint4 a = rec[3 * s], b = rec[3 * s + 1], c = rec[3 * s + 2];
        p = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
        v = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
        o = c.x; no = (int)c.y; key = keys_own[e - low];
    } else {
        const size_t hidx = e < low ? e : e - nown_;
        const uint4 a = halo[2 * hidx], b = halo[2 * hidx + 1];
        p = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
        v = make_float4(0.f, 0.f, 0.f, __uint_as_float(b.x));
        o = b.y; key = b.z;
    }
    posh[e] = p; velm[e] = v; orig[e] = o; nown[e] = no; keys_ext[e] = key;
    posm[e] = make_float4(p.x, p.y, p.z, v.w);
    posc[e] = make_float4(p.x, p.y, p.z, sph_keep_threshold(p.w));
    const int shift = 3 * (10 - g->bits);
    const uint32_t ck = key >> shift;
    if (e == 0 || (key_of(e - 1) >> shift) != ck) cell_start[ck] = (uint32_t)e;
    if (e == next - 1 || (key_of(e + 1) >> shift) != ck) cell_end[ck] = (uint32_t)(e + 1);
    atomicMax(&cell_hmax[ck], __float_as_uint(p.w));
}

__global__ void __launch_bounds__(256) k_gather_f32(const float* __restrict__ src, const uint32_t* __restrict__ list, int n,
                                                    float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = src[list[k]];
}

// first / last SPH_TOP_LEAF particles of the rank: bnd[(side * SPH_TOP_LEAF + slot) * 2 + {0,1}] = posh, velm; the last
// ones are right-aligned (slot = SPH_TOP_LEAF - (n - i))
__global__ void k_boundary(const float4* __restrict__ posh, const float4* __restrict__ velm, int n, float4* __restrict__ bnd) {
    const int k = threadIdx.x;   // 2 * SPH_TOP_LEAF threads
    const int side = k / SPH_TOP_LEAF, slot = k % SPH_TOP_LEAF;
    const int i = side == 0 ? slot : n - SPH_TOP_LEAF + slot;
    if (i < 0 || i >= n) return;
    bnd[2 * (size_t)k] = posh[i];
    bnd[2 * (size_t)k + 1] = velm[i];
}


// ---- locally essential tree (LET): which of this rank's finished walk nodes the other ranks can reach ------------------------
// A walk reads node k only after a target of the walking rank has opened k's parent p (rejected it: r_sq <= T_p, kernels_tree.cu).
// With B_r = the box of the positions rank r walks (its own targets plus the <= 31 companions that fill its first and last
// 32-slot group), "some target of r opens p" implies dist^2(cm_p, B_r) <= T_p: the sender evaluates this superset test for every
// one of its nodes against every other rank's box and ships only those records instead of all-gathering the whole node array.
// Nodes inside buckets are unreachable (mask 0); nodes hanging below a straddling (top) node are needed by everybody.

// box of posm[lo, hi) into 8 ordered uints (lo.xyz, -, hi.xyz, -), pre-set to (0xffffffff x4, 0 x4)
__global__ void __launch_bounds__(256) k_let_box(const float4* __restrict__ posm, int lo, int hi, uint32_t* __restrict__ box) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        const float4 p = posm[i];
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint32_t a = __reduce_min_sync(FULL, f2ord(mn[k])), b = __reduce_max_sync(FULL, f2ord(mx[k]));
        if ((threadIdx.x & 31) == 0) { atomicMin(&box[k], a); atomicMax(&box[4 + k], b); }
    }
}

// mask[u] = ranks that need own node u (u enumerates the internal nodes g0..g1-1, then the leaves of slots g0..g1-1);
// cnt[r] += number of nodes rank r needs
__global__ void __launch_bounds__(256) k_let_mask(const int2* __restrict__ range, const int32_t* __restrict__ parent,
                                                  const float4* __restrict__ packed, int n, int g0, int g1, int leaf_max,
                                                  const uint32_t* __restrict__ boxes, int world, int me, uint32_t* __restrict__ mask,
                                                  uint32_t* __restrict__ cnt) {
    __shared__ float sbox[SPH_MAX_RANKS][6];
    if (threadIdx.x < world * 6) {
        const int r = threadIdx.x / 6, k = threadIdx.x % 6;
        sbox[r][k] = ord2f(boxes[8 * r + (k < 3 ? k : k + 1)]);
    }
    __syncthreads();
    const int u = blockIdx.x * blockDim.x + threadIdx.x, own = g1 - g0;
    uint32_t m = 0u;
    if (u < 2 * own) {
        const int k = u < own ? g0 + u : n - 1 + g0 + (u - own);
        if (!(u < own && k >= n - 1)) {
            const int2 rg = range[k];
            const bool top = rg.x < g0 || rg.y >= g1;            // straddles a rank boundary: finished by k_top_tree on every rank
            const int p = parent[k];
            const uint32_t all = (world >= 32 ? 0xffffffffu : ((1u << world) - 1u)) & ~(1u << me);
            if (!top) {
                if (p < 0) m = all;                                // root, or the parent is built elsewhere: below a top node
                else {
                    const int2 prg = range[p];
                    if (prg.y - prg.x + 1 > leaf_max) {            // else: strictly inside a bucket, unreachable
                        if (prg.x < g0 || prg.y >= g1) m = all;    // the parent is a top node
                        else {
                            const float4 N = packed[2 * (size_t)p];   // (cm, T)
                            for (int r = 0; r < world; r++) {
                                if (r == me) continue;
                                const float dx = fmaxf(fmaxf(sbox[r][0] - N.x, N.x - sbox[r][3]), 0.f);
                                const float dy = fmaxf(fmaxf(sbox[r][1] - N.y, N.y - sbox[r][4]), 0.f);
                                const float dz = fmaxf(fmaxf(sbox[r][2] - N.z, N.z - sbox[r][5]), 0.f);
                                // an empty box (a rank without targets) has lo = +inf: d = inf, never needed
                                if ((dx * dx + dy * dy + dz * dz) * 0.99999f <= N.w) m |= 1u << r;
                            }
                        }
                    }
                }
            }
        }
        mask[u] = m;
    }
    for (int r = 0; r < world; r++) {
        const unsigned b = __ballot_sync(FULL, (m >> r) & 1u);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&cnt[r], (uint32_t)__popc(b));
    }
}

// records (cm,T | M,a,b,Bmax^2 | id) of the needed nodes, grouped by destination: segment of rank r starts at soff[r];
// cursor[r] counts what has been placed (the order inside a segment is irrelevant: the receiver scatters by id)
__global__ void __launch_bounds__(256) k_let_pack(const uint32_t* __restrict__ mask, const float4* __restrict__ packed, int n, int g0, int g1,
                                                  int world, const uint32_t* __restrict__ soff, uint32_t* __restrict__ cursor,
                                                  float4* __restrict__ out) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, own = g1 - g0;
    const int lane = threadIdx.x & 31;
    uint32_t m = u < 2 * own ? mask[u] : 0u;
    const int k = u < own ? g0 + u : n - 1 + g0 + (u - own);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (m) { a = packed[2 * (size_t)k]; b = packed[2 * (size_t)k + 1]; }
    for (int r = 0; r < world; r++) {
        const unsigned bal = __ballot_sync(FULL, (m >> r) & 1u);
        if (!bal) continue;
        uint32_t base = 0u;
        if (lane == 0) base = atomicAdd(&cursor[r], (uint32_t)__popc(bal));
        base = __shfl_sync(FULL, base, 0);
        if ((m >> r) & 1u) {
            float4* o = out + 3 * ((size_t)soff[r] + base + __popc(bal & ((1u << lane) - 1u)));
            o[0] = a; o[1] = b; o[2] = make_float4(__int_as_float(k), 0.f, 0.f, 0.f);
        }
    }
}

__global__ void __launch_bounds__(256) k_let_scatter(const float4* __restrict__ rec, int nrec, long long n_nodes, float4* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    const int k = __float_as_int(rec[3 * (size_t)i + 2].x);
    SPH_DBG_IDX(k, 2 * (long long)n_nodes);
    packed[2 * (size_t)k] = rec[3 * (size_t)i];
    packed[2 * (size_t)k + 1] = rec[3 * (size_t)i + 1];
}

// ---- download: results travel back to the rank that holds the particle's body-order slice
__global__ void __launch_bounds__(256) k_body_dest(const uint32_t* __restrict__ orig, int n, int64_t chunk, uint8_t* __restrict__ dest) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dest[i] = (uint8_t)((int64_t)orig[i] / chunk);
}

// result record, 6 x 16 B: [0] pos, h  [1] vel, m  [2] gradP, rho  [3] gradPhi, Phi  [4] P, own-support count, neighbor count,
// numParticles  [5] numApprox, body index
__global__ void __launch_bounds__(256) k_result_pack(const uint32_t* __restrict__ perm, int n, const float4* __restrict__ posh,
                                                     const float4* __restrict__ velm, const uint32_t* __restrict__ orig,
                                                     const float* __restrict__ rho, const float* __restrict__ press,
                                                     const float4* __restrict__ gradp, const float4* __restrict__ grav,
                                                     const int32_t* __restrict__ nown, const int32_t* __restrict__ ncount,
                                                     const int32_t* __restrict__ npart, const int32_t* __restrict__ napprox,
                                                     float4* __restrict__ rec) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t i = perm[k];
    float4* r = rec + 6 * (size_t)k;
    const float4 gp = gradp[i];
    r[0] = posh[i]; r[1] = velm[i];
    r[2] = make_float4(gp.x, gp.y, gp.z, rho[i]);
    r[3] = grav[i];
    r[4] = make_float4(press[i], __int_as_float(nown[i]), __int_as_float(ncount[i]), __int_as_float(npart[i]));
    r[5] = make_float4(__int_as_float(napprox[i]), __uint_as_float(orig[i]), 0.f, 0.f);
}

// one field of the received records into a dense staging array in body order (element b = body index - body0)
__global__ void __launch_bounds__(256) k_result_field(const float4* __restrict__ rec, int n, int field, int64_t body0, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4* r = rec + 6 * (size_t)k;
    const size_t b = (size_t)((int64_t)__float_as_uint(r[5].y) - body0);
    switch (field) {
        case SPH_FIELD_TRANSLATION: { const float4 p = r[0]; out[3 * b] = p.x; out[3 * b + 1] = p.y; out[3 * b + 2] = p.z; break; }
        case SPH_FIELD_VELOCITY: { const float4 v = r[1]; out[3 * b] = v.x; out[3 * b + 1] = v.y; out[3 * b + 2] = v.z; break; }
        case SPH_FIELD_MASS: out[b] = r[1].w; break;
        case SPH_FIELD_SMOOTHING: out[2 * b] = r[0].w; out[2 * b + 1] = r[4].y; break;
        case SPH_FIELD_COUNT_: {   // the full sph_ParticleSmoothing record (ParticleSmoothing.cs:16-31), 7 words
            const float h = r[0].w;
            float* o = out + 7 * b;
            o[0] = h; o[1] = 2.0f * h; o[2] = 0.f; o[3] = 0.f; o[4] = 0.f; o[5] = 2.0f * h; o[6] = r[4].y;
            break;
        }
        case SPH_FIELD_DENSITY: out[b] = r[2].w; break;
        case SPH_FIELD_PRESSURE: out[b] = r[4].x; break;
        case SPH_FIELD_PRESSURE_GRAD: { const float4 q = r[2]; out[3 * b] = q.x; out[3 * b + 1] = q.y; out[3 * b + 2] = q.z; break; }
        case SPH_FIELD_GRAVITY: {
            const float4 q = r[3];
            out[6 * b] = q.x; out[6 * b + 1] = q.y; out[6 * b + 2] = q.z; out[6 * b + 3] = q.w;
            out[6 * b + 4] = r[4].w; out[6 * b + 5] = r[5].x;
            break;
        }
        case SPH_FIELD_NEIGHBOR_COUNT: out[b] = r[4].z; break;
    }
}

// mass range of the uploaded particles (equal-mass detection without a host pass): ordered-uint min / max
__global__ void __launch_bounds__(256) k_mass_range(const float4* __restrict__ velm, int n, uint32_t* __restrict__ mm) {
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t u = f2ord(velm[i].w);
        lo = min(lo, u); hi = max(hi, u);
    }
    lo = __reduce_min_sync(FULL, lo); hi = __reduce_max_sync(FULL, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}

// in-process transport: buf[i] = op over the ranks' copies scratch[q][i]
__global__ void __launch_bounds__(256) k_reduce_ranks(const void* __restrict__ scratch, int world, size_t count, int dtype, int op,
                                                      void* __restrict__ buf) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (dtype == GR_U32) {
        const uint32_t* s = (const uint32_t*)scratch;
        uint32_t v = s[i];
        for (int q = 1; q < world; q++) {
            const uint32_t x = s[(size_t)q * count + i];
            v = op == GR_SUM ? v + x : op == GR_MIN ? min(v, x) : max(v, x);
        }
        ((uint32_t*)buf)[i] = v;
    } else if (dtype == GR_U64) {
        const unsigned long long* s = (const unsigned long long*)scratch;
        unsigned long long v = s[i];
        for (int q = 1; q < world; q++) {
            const unsigned long long x = s[(size_t)q * count + i];
            v = op == GR_SUM ? v + x : op == GR_MIN ? min(v, x) : max(v, x);
        }
        ((unsigned long long*)buf)[i] = v;
    } else {
        const double* s = (const double*)scratch;
        double v = s[i];
        for (int q = 1; q < world; q++) {
            const double x = s[(size_t)q * count + i];
            v = op == GR_SUM ? v + x : op == GR_MIN ? fmin(v, x) : fmax(v, x);
        }
        ((double*)buf)[i] = v;
    }
}

}  // namespace

// ---- launchers (stream = the rank's context stream) -------------------------------------------------------------------
#define GL(c) SPH_LAUNCH_CHECK(c)

int grk_bin_hist(sphb200_ctx* c, const uint32_t* keys, const int32_t* ncount, const int32_t* npart, const int32_t* napprox, int n,
                 uint32_t* hist) {
    SPH_CK(c, cudaMemsetAsync(hist, 0, 2 * (size_t)SPH_NBINS * sizeof(uint32_t), c->stream));
    if (n > 0) { k_bin_hist<<<sph_div_up(n, 256), 256, 0, c->stream>>>(keys, ncount, npart, napprox, n, c->grid_d, hist); GL(c); }
    return SPH_OK;
}
int grk_splitters(sphb200_ctx* c, const uint32_t* hist, int world, int64_t* split, unsigned long long* part, uint32_t* bmask) {
    k_bin_partials<<<SPL_NRUN, SPL_RUN, 0, c->stream>>>(hist, part); GL(c);
    k_splitters<<<1, SPL_RUN, 0, c->stream>>>(hist, part, world, split); GL(c);
    k_bin_owner_mask<<<SPH_NBINS / 256, 256, 0, c->stream>>>(c->grid_d, split, world, bmask); GL(c);
    return SPH_OK;
}
int grk_dest(sphb200_ctx* c, const uint32_t* keys, int n, const int64_t* split, int world, uint8_t* dest) {
    if (n > 0) { k_dest<<<sph_div_up(n, 256), 256, 0, c->stream>>>(keys, n, c->grid_d, split, world, dest); GL(c); }
    return SPH_OK;
}
int grk_mig_pack(sphb200_ctx* c, const float4* posh, const float4* velm, const uint32_t* orig, const int32_t* nown, const uint32_t* keys,
                 const uint32_t* perm, int n, uint4* rec) {
    if (n > 0) { k_mig_pack<<<sph_div_up(n, 256), 256, 0, c->stream>>>(posh, velm, orig, nown, keys, perm, n, rec); GL(c); }
    return SPH_OK;
}
int grk_mig_keys(sphb200_ctx* c, const uint4* rec, int n, uint32_t* keys) {
    if (n > 0) { k_mig_keys<<<sph_div_up(n, 256), 256, 0, c->stream>>>(rec, n, keys); GL(c); }
    return SPH_OK;
}
// halo lists of the sorted own particles: mask -> per-destination stable compaction; total[q] = particles for rank q
int grk_halo_lists(sphb200_ctx* c, const uint32_t* keys, int n, const int64_t* split, int world, int me, const uint32_t* bmask, uint32_t* mask,
                   uint32_t* cnt, uint32_t* total, uint32_t* list) {
    SPH_CK(c, cudaMemsetAsync(total, 0, 256 * sizeof(uint32_t), c->stream));
    if (n <= 0 || world <= 1) return SPH_OK;
    const int ntiles = sph_div_up(n, HC_TILE);
    k_halo_mask<<<sph_div_up(n, 256), 256, 0, c->stream>>>(keys, n, c->grid_d, split, world, me, bmask, mask); GL(c);
    k_halo_count<<<ntiles, HC_THREADS, 0, c->stream>>>(mask, n, world, ntiles, cnt); GL(c);
    int rc = sph_launch_rowscan(c, cnt, ntiles, world, total, c->stream);
    if (rc) return rc;
    k_halo_scatter<<<ntiles, HC_THREADS, 0, c->stream>>>(mask, n, world, ntiles, cnt, total, list); GL(c);
    return SPH_OK;
}
size_t grk_halo_cnt_words(int64_t cap) { return (size_t)SPH_MAX_RANKS * (size_t)sph_div_up(cap, HC_TILE) + 256; }

int grk_halo_pack(sphb200_ctx* c, const uint4* rec, const uint32_t* idx, const uint32_t* keys, const uint32_t* list, int n, uint4* out) {
    if (n > 0) { k_halo_pack<<<sph_div_up(n, 256), 256, 0, c->stream>>>(rec, idx, keys, list, n, out); GL(c); }
    return SPH_OK;
}
int grk_assemble_ext(sphb200_ctx* c, const uint4* rec, const uint32_t* idx, const uint32_t* keys_own, const uint4* halo, int low, int nown,
                     int next, uint32_t* keys_ext, size_t ncell) {
    SPH_CK(c, cudaMemsetAsync(c->cell_start, 0, ncell * sizeof(uint32_t), c->stream));
    SPH_CK(c, cudaMemsetAsync(c->cell_end, 0, ncell * sizeof(uint32_t), c->stream));
    SPH_CK(c, cudaMemsetAsync(c->cell_hmax, 0, ncell * sizeof(uint32_t), c->stream));
    if (next > 0) {
        k_assemble_ext<<<sph_div_up(next, 256), 256, 0, c->stream>>>(rec, idx, keys_own, halo, low, nown, next, c->grid_d, c->posh[0], c->velm[0],
                                                                    c->orig[0], c->nown, keys_ext, c->posm, c->posc, c->cell_start,
                                                                    c->cell_end, c->cell_hmax);
        GL(c);
    }
    return SPH_OK;
}
int grk_sorted_sources(sphb200_ctx* c, const uint4* rec, const uint32_t* idx, const uint32_t* keys, int n, float4* posm, uint32_t* keys_out,
                       float4* tposh, float4* tvelm) {
    if (n > 0) { k_sorted_sources<<<sph_div_up(n, 256), 256, 0, c->stream>>>(rec, idx, keys, n, posm, keys_out, tposh, tvelm); GL(c); }
    return SPH_OK;
}
int grk_gather_f32(sphb200_ctx* c, const float* src, const uint32_t* list, int n, float* out) {
    if (n > 0) { k_gather_f32<<<sph_div_up(n, 256), 256, 0, c->stream>>>(src, list, n, out); GL(c); }
    return SPH_OK;
}
int grk_boundary(sphb200_ctx* c, const float4* posh, const float4* velm, int n, float4* bnd) {
    SPH_CK(c, cudaMemsetAsync(bnd, 0, 2 * SPH_TOP_LEAF * 2 * sizeof(float4), c->stream));
    if (n > 0) { k_boundary<<<1, 2 * SPH_TOP_LEAF, 0, c->stream>>>(posh, velm, n, bnd); GL(c); }
    return SPH_OK;
}

// box[8] (ordered uints) of the positions posm[lo, hi); an empty range leaves lo = +inf
int grk_let_box(sphb200_ctx* c, const float4* posm, int lo, int hi, uint32_t* box) {
    const uint32_t init[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u};
    SPH_CK(c, cudaMemcpyAsync(box, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    if (hi > lo) { k_let_box<<<min(sph_div_up(hi - lo, 256), c->sm_count * 4), 256, 0, c->stream>>>(posm, lo, hi, box); GL(c); }
    return SPH_OK;
}
int grk_let_mask(sphb200_ctx* c, const uint32_t* boxes, int world, int me, uint32_t* mask, uint32_t* cnt) {
    const int n = (int)c->tree_n, g0 = (int)c->tree_g0, g1 = (int)c->tree_g1;
    SPH_CK(c, cudaMemsetAsync(cnt, 0, (size_t)world * sizeof(uint32_t), c->stream));
    if (g1 > g0) {
        k_let_mask<<<sph_div_up(2 * (int64_t)(g1 - g0), 256), 256, 0, c->stream>>>(c->range, c->parent, c->packed, n, g0, g1, c->p.leaf_max, boxes, world, me, mask, cnt);
        GL(c);
    }
    return SPH_OK;
}
int grk_let_pack(sphb200_ctx* c, const uint32_t* mask, int world, const uint32_t* soff, uint32_t* cursor, float4* out) {
    const int n = (int)c->tree_n, g0 = (int)c->tree_g0, g1 = (int)c->tree_g1;
    SPH_CK(c, cudaMemsetAsync(cursor, 0, (size_t)world * sizeof(uint32_t), c->stream));
    if (g1 > g0) { k_let_pack<<<sph_div_up(2 * (int64_t)(g1 - g0), 256), 256, 0, c->stream>>>(mask, c->packed, n, g0, g1, world, soff, cursor, out); GL(c); }
    return SPH_OK;
}
int grk_let_scatter(sphb200_ctx* c, const float4* rec, int64_t nrec) {
    if (nrec > 0) { k_let_scatter<<<sph_div_up(nrec, 256), 256, 0, c->stream>>>(rec, (int)nrec, (long long)c->tree_n, c->packed); GL(c); }
    return SPH_OK;
}
int grk_body_dest(sphb200_ctx* c, const uint32_t* orig, int n, int64_t chunk, uint8_t* dest) {
    if (n > 0) { k_body_dest<<<sph_div_up(n, 256), 256, 0, c->stream>>>(orig, n, chunk, dest); GL(c); }
    return SPH_OK;
}
int grk_result_pack(sphb200_ctx* c, const uint32_t* perm, int n, int own0, float4* rec) {
    if (n > 0) {
        k_result_pack<<<sph_div_up(n, 256), 256, 0, c->stream>>>(perm, n, c->posh[0] + own0, c->velm[0] + own0, c->orig[0] + own0, c->rho + own0,
                                                                c->press + own0, c->gradp + own0, c->grav + own0, c->nown + own0,
                                                                c->ncount + own0, c->npart + own0, c->napprox + own0, rec);
        GL(c);
    }
    return SPH_OK;
}
int grk_result_field(sphb200_ctx* c, const float4* rec, int n, int field, int64_t body0, float* out) {
    if (n > 0) { k_result_field<<<sph_div_up(n, 256), 256, 0, c->stream>>>(rec, n, field, body0, out); GL(c); }
    return SPH_OK;
}
int grk_mass_range(sphb200_ctx* c, const float4* velm, int n, uint32_t* mm) {
    if (n > 0) { k_mass_range<<<min(sph_div_up(n, 256), c->sm_count * 4), 256, 0, c->stream>>>(velm, n, mm); GL(c); }
    return SPH_OK;
}
int grk_reduce_ranks(sphb200_ctx* c, const void* scratch, int world, size_t count, int dtype, int op, void* buf) {
    if (count > 0) { k_reduce_ranks<<<sph_div_up((int64_t)count, 256), 256, 0, c->stream>>>(scratch, world, count, dtype, op, buf); GL(c); }
    return SPH_OK;
}
