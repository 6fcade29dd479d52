// ctx.cuh -- internal context + helpers of libsphb200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/sphb200.h"

// ---- SPH_DEBUG_BOUNDS build (make debug -> libsphb200_dbg.so): stands in for compute-sanitizer, which the GPU pool refuses.
//   * every data-dependent shared-memory / row / stack / list index in the kernels is asserted (SPH_DBG_IDX): a violation
//     prints file:line, the index and its limit, and traps -- the next synchronising ABI call fails;
//   * every device allocation of the library carries 256-byte guard zones on both sides, filled with 0xA5;
//     sphb200_debug_check_guards counts the guard bytes that no longer hold the pattern (out-of-bounds WRITES of any kernel).
// In the release build SPH_DBG_IDX compiles to nothing and allocations are plain cudaMalloc / cudaFree.
#ifdef SPH_DEBUG_BOUNDS
#include <stdio.h>
#define SPH_DBG_IDX(i, n)                                                                                                     \
    do {                                                                                                                      \
        if (!((long long)(i) >= 0 && (long long)(i) < (long long)(n))) {                                                      \
            printf("SPH_DEBUG_BOUNDS %s:%d: index %lld outside [0, %lld) (block %d thread %d)\n", __FILE__, __LINE__,           \
                   (long long)(i), (long long)(n), (int)blockIdx.x, (int)threadIdx.x);                                        \
            __trap();                                                                                                         \
        }                                                                                                                     \
    } while (0)
#else
#define SPH_DBG_IDX(i, n) ((void)0)
#endif
cudaError_t sph_dev_malloc(void** p, size_t bytes);   // abi.cu
cudaError_t sph_dev_free(void* p);
int sph_debug_guard_errors(long long* bad_bytes, long long* allocations);

#define SPH_MAX_PASSES 24
#define SPH_RR_TABLE 4096          // radius-ratio table entries (own-support count 1..4095)
#define SPH_MAX_RANKS 32           // ranks of one group (halo masks are 32-bit)
#define SPH_TOP_LEAF 64            // >= the largest leaf_max: boundary particles published per rank side
#define SPH_TOP_CAP 512            // per-rank capacity of the top-tree lists (nodes that straddle a rank boundary)

// ERR_NEIGHBOR_OVERFLOW describes the lists in memory (re-armed by every neighbor pass); ERR_OVERFLOW_EVER keeps the largest count
// any pass since the upload has seen
enum { ERR_NEIGHBOR_OVERFLOW = 0, ERR_TREE_STACK = 1, ERR_TOP_TREE = 2, ERR_OVERFLOW_EVER = 3, ERR_SLOTS = 4 };

// Top-tree records of the distributed LBVH build (kernels_tree.cu / kernels_group.cu).  A node whose particle range crosses
// a rank boundary cannot be finished by one rank: its owner emits its topology, the owners of its finished children emit
// their moments, and every rank then completes the few straddling nodes redundantly (bit-identical on all ranks).
struct TopNode { int id, a, b, lo, hi, pad0, pad1, pad2; };        // a/b = children (internal) ; lo..hi = particle range
struct FrontNode { int id, pad0, pad1, pad2; float4 mom, lo, hi; };   // a finished node whose parent is a top node

struct sphb200_ctx {
    sph_Params p{};
    int device = 0;
    int sm_count = 148;
    int64_t cap = 0, n = 0;      // resident slots: capacity, in use
    int64_t cap_rows = 0;        // slots that can be targets (neighbor rows, all-pairs partial sums, staging)
    cudaStream_t own_stream = nullptr, stream = nullptr;

    // resident SoA state, sorted (Morton) order after build_neighbors; ping-pong through the permute
    float4* posh[2] = {nullptr, nullptr};  // x,y,z,h
    float4* velm[2] = {nullptr, nullptr};  // vx,vy,vz,m
    uint32_t* orig[2] = {nullptr, nullptr};  // sorted slot -> body index
    int cur = 0;
    float4* posm = nullptr;                // x,y,z,m (gravity sources)
    float4* posc = nullptr;                // x,y,z,C(h): neighbor-test records (sph_keep_threshold)
    unsigned int* chunk_counter = nullptr; // work counter of k_cell_neighbors

    uint32_t* keys[2] = {nullptr, nullptr};
    uint32_t* idx[2] = {nullptr, nullptr};
    uint32_t* skeys = nullptr;               // sorted keys of the resident slots (single handle: keys[1]; group rank: extended set)
    uint32_t* sort_hist = nullptr;   // radix sort: digit counts [256][tiles] + totals [256]
    uint32_t* cell_start = nullptr;
    uint32_t* cell_end = nullptr;
    uint32_t* cell_hmax = nullptr;  // max h per cell (fp32 bit pattern)
    int grid_bits_max = 0;
    size_t ncell_max = 0;

    uint32_t* nlist = nullptr;   // [rows][max_neighbors] sorted-slot indices; row of slot t = t - row_base
    int64_t row_base = 0;        // group ranks keep rows for their own slots only
    int32_t* ncount = nullptr;   // symmetric neighbor count
    int32_t* nown = nullptr;     // own-support count (KernelThis.w > 0)
    float* rho = nullptr;
    float* press = nullptr;
    float* cvol = nullptr;       // (m/rho)*P
    float4* gradp = nullptr;
    float4* grav = nullptr;
    int32_t* npart = nullptr;
    int32_t* napprox = nullptr;
    float4* tbox = nullptr;      // all-pairs: bounding box (lo, hi) of every 256-source tile
    bool equal_mass = false;     // every uploaded particle has the same mass (ParticleAuthoring.cs:208)
    float common_mass = 0.f;
    float h_bound = INFINITY;    // host-side upper bound of every resident h (upload value x the controller's largest growth per update):
                                 // below 1e5 the literal-kernel neighbor pass (k_neighbors_density) cannot be needed and is not launched
    float4* gpart = nullptr;     // all-pairs partial sums [splits][n]
    size_t gpart_entries = 0;    // capacity of gpart: 32 splits of the target rows where that stays under 1 GB, never fewer than 8
    float4* gsrc = nullptr;      // gravity sources (x,y,z,m) in GLOBAL sorted order: posm for a single handle, the all-gathered
    int64_t gsrc_n = 0;          // array for a group rank
    // LBVH addressing: the tree spans tree_n global slots; this context builds the nodes of slots [tree_g0, tree_g1) and its
    // resident slot of global slot s is s + tree_off
    uint32_t* tkeys = nullptr;
    int64_t tree_n = 0, tree_g0 = 0, tree_g1 = 0, tree_off = 0;
    // particle records the LBVH build reads: slot s of the tree at tree_posh[s + tree_src_off] (null: the resident posh / velm at
    // s + tree_off).  A group rank builds from the sorted own records before its extended resident set is assembled.
    const float4* tree_posh = nullptr;
    const float4* tree_velm = nullptr;
    int64_t tree_src_off = 0;
    TopNode* top_nodes = nullptr;     // [world][SPH_TOP_CAP] (own segment filled by the build, the rest all-gathered)
    FrontNode* front_nodes = nullptr; // [world][SPH_TOP_CAP]
    int32_t* top_counts = nullptr;    // [world][4]: top nodes, frontier nodes
    int top_rank = 0;                 // own segment

    // LBVH (2n-1 nodes)
    int2* child = nullptr;
    int2* range = nullptr;
    int32_t* parent = nullptr;
    int32_t* flag = nullptr;
    float4* mom = nullptr;
    float4* nlo = nullptr;       // xyz = box min, w = first (int bits)
    float4* nhi = nullptr;       // xyz = box max, w = last  (int bits)
    float4* packed = nullptr;    // walk nodes, 2 float4 per node: (cm, M), (Bmax^2, a, b, -)

    uint32_t* bounds = nullptr;  // [0..6] ordered-uint floats: min xyz, max xyz, hmax; [8..9] u64 sum of h bit patterns
    sph_GridParams* grid_d = nullptr;
    sph_GridParams grid_h{};
    int32_t* err_d = nullptr;
    int32_t* err_h = nullptr;    // pinned
    float* rr_table = nullptr;
    double* diag_d = nullptr;

    void* stage_d = nullptr;     // upload/download staging (device)
    void* stage_h = nullptr;     // pinned host staging
    size_t stage_bytes = 0;
    void* scratch_d = nullptr;   // grow-only device scratch of the list/record downloads (no cudaMalloc per call once sized)
    size_t scratch_bytes = 0;

    bool resident = false, lists_valid = false, pressure_valid = false, gravity_valid = false, tree_valid = false;
    bool h_updated = false;      // bounds/grid params computed for the current positions and h
    bool sorted_valid = false;   // sort + cell table match the current positions and h
    bool lists_fresh = false;    // neighbor lists/density belong to the current positions and h
    bool nown_aligned = false;   // nown[] is index-aligned with posh[cur] (true after upload and after the density pass; a sort
                                 // permutes the slots but not nown, which the next density pass rewrites)
    // overlapped LBVH build (sphb200_prepare_gravity): built on aux_stream while the neighbor pass runs on `stream`
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // recorded behind the pressure pass: rho, P, grad P, the neighbor counts and h are final from here until the next smoothing
    // update / neighbor pass, so their downloads run on the auxiliary stream beside whatever the step still has to do (gravity)
    cudaEvent_t ev_sph = nullptr;
    bool sph_ready = false;
    bool tree_hint = false, tree_fresh = false, tree_join_pending = false;
    float hint_dt = 0.f, tree_dt = 0.f;
    int64_t t0 = 0, t1 = -1;
    int64_t launches = 0;
    float last_dt = 0.f;

    // timing
    bool timing = false;
    int npass = 0;
    const char* pass_name[SPH_MAX_PASSES];
    cudaEvent_t ev[SPH_MAX_PASSES + 1];
    bool ev_created = false;

    std::string err;
};

#define SPH_CK(ctx, call)                                                                            \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return SPH_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

#define SPH_LAUNCH_CHECK(ctx)                                                                        \
    do {                                                                                             \
        (ctx)->launches++;                                                                           \
        cudaError_t e__ = cudaGetLastError();                                                        \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + std::to_string(__LINE__); \
            return SPH_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

static inline int sph_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- device helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t compact10(uint32_t v) {
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030C30C3u;
    v = (v | (v >> 4)) & 0x0300F00Fu;
    v = (v | (v >> 8)) & 0x030000FFu;
    v = (v | (v >> 16)) & 0x3ffu;
    return v;
}
// exact (non-contracted) squared distance in the reference's order: dx*dx + dy*dy + dz*dz
__device__ __forceinline__ float dot3_rn(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// Neighbor keep threshold C(h) (kernels_neighbors.cu): the pair (i,j) is kept iff d2 < max(C(h_i), C(h_j)), where
// C(h) = min( ((h*h)*2)*2 , t(h) ) and t(h) is the smallest float x with fsqrt_rn(x) >= 2h -- the reference's
// "d2 < size*size*Kappa*Kappa" (SplineKernel.cs:47-53) and "Kernel(r,h) > 0 <=> r < 2h" (:62) without the sqrt.
// t(h): the smallest float x with fsqrt_rn(x) >= 2h, i.e. fsqrt_rn(d2) < 2h <=> d2 < t(h) -- "inside the support of h"
// (Kernel(r,h) > 0 <=> r < 2h, SplineKernel.cs:62) decided on the squared distance.
__device__ __forceinline__ float sph_own_threshold(float h) {
    if (!(h > 0.0f)) return 0.0f;   // h <= 0 or NaN (invalid input): empty support, and the ulp search below must not run
    const float c = __fmul_rn(h, 2.0f);
    uint32_t u = __float_as_uint(__fmul_rn(c, c));
    while (u > 0u && __fsqrt_rn(__uint_as_float(u - 1u)) >= c) --u;
    while (__fsqrt_rn(__uint_as_float(u)) < c) ++u;
    return __uint_as_float(u);
}
__device__ __forceinline__ float sph_keep_threshold(float h) {
    if (!(h > 0.0f)) return 0.0f;
    const float a = __fmul_rn(__fmul_rn(__fmul_rn(h, h), 2.0f), 2.0f);
    return fminf(a, sph_own_threshold(h));
}

int sph_grid_bits_for(const sph_Params& p, int64_t n);
int sph_validate_params(const sph_Params& p, int64_t capacity, std::string& err);
int sph_ctx_create(const sph_Params& p, int device, int64_t slots, int64_t rows, int64_t tree, int64_t total, bool pingpong,
                   sphb200_ctx** out, std::string& err);

// ---- kernel launchers (one translation unit each) -----------------------------------------------------------
int sph_launch_smoothing_bounds(sphb200_ctx* c, bool update_h);
int sph_launch_sort_and_cells(sphb200_ctx* c);
int sph_launch_bounds_range(sphb200_ctx* c, float4* posh, const int32_t* nown, int n, bool update_h);
int sph_launch_grid_setup(sphb200_ctx* c, int64_t n_total);
int sph_launch_keys(sphb200_ctx* c, const float4* posh, int n, uint32_t* keys);
int sph_launch_rowscan(sphb200_ctx* c, uint32_t* rows_d, int nblocks, int rows, uint32_t* totals, cudaStream_t stream);
int sph_launch_neighbors_density(sphb200_ctx* c);
int sph_launch_pressure(sphb200_ctx* c);
int sph_launch_gravity_near(sphb200_ctx* c, int impl);
int sph_launch_integrate(sphb200_ctx* c, float dt);
int sph_launch_gravity_allpairs(sphb200_ctx* c);
int sph_launch_tree_build(sphb200_ctx* c, float dt, cudaStream_t stream);
int sph_launch_tree_walk(sphb200_ctx* c);
int sph_launch_top_tree(sphb200_ctx* c, int world, const float4* bnd, const int64_t* g0_d, float dt);
int sph_launch_diagnostics(sphb200_ctx* c, double* out12);
int sph_launch_diagnostics_range(sphb200_ctx* c, int off, int n);
int sph_launch_field_stats_range(sphb200_ctx* c, int off, int n);
void sph_finish_field_stats(const double* sums4, const uint32_t* mm8, int64_t n, double* out12);
int sph_launch_pack_upload(sphb200_ctx* c, int64_t n, int has_nown, uint32_t orig0);
int sph_upload_core(sphb200_ctx* c, int64_t n, uint32_t orig0, const void* pos, int pos_stride, const void* vel, int vel_stride,
                    const void* mass, int mass_stride, const void* smoothing, int smoothing_stride);
int sph_upload_finish(sphb200_ctx* c);
int sph_launch_unpack_field(sphb200_ctx* c, int field, int* elem_bytes);
int sph_launch_neighbor_rows_sorted(sphb200_ctx* c, int32_t* rows_d);
int sph_launch_interactions(sphb200_ctx* c, int64_t total, const int64_t* offsets_d, const int32_t* nbr_d, sph_ParticleInteraction* out_d);
int sph_fp32_peak(sphb200_ctx* c, double* tflops);
int sph_snapshot_write(const char* path, int64_t n_total, int64_t body0, int64_t count, int64_t steps, const float* pos, const float* vel,
                       const float* mass, const float* h, const int32_t* nown, std::string& err);
int sph_snapshot_read(const char* path, int64_t want0, int64_t want, int64_t* n_total, std::vector<float>& pos, std::vector<float>& vel,
                      std::vector<float>& mass, std::vector<sph_ParticleSmoothing>& sm, std::string& err);
size_t sph_sort_hist_words(int64_t cap);
int sph_launch_radix_sort(sphb200_ctx* c, int n, cudaStream_t stream);
int sph_launch_digit_pass(sphb200_ctx* c, const uint32_t* keys_in, const uint32_t* vals_in, const uint8_t* bucket, int n, int shift,
                          uint32_t* keys_out, uint32_t* vals_out, uint32_t* total_out, cudaStream_t stream);
