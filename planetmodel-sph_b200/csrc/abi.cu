// abi.cu -- the C ABI of libsphb200 (see include/sphb200.h for the contract and the reference citations).
#include "ctx.cuh"
#include <stdio.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <new>

static std::string g_create_err;

#define ARG_CHECK(c, cond, msg)                \
    do {                                       \
        if (!(cond)) {                         \
            (c)->err = (msg);                  \
            return SPH_ERR_INVALID_ARG;        \
        }                                      \
    } while (0)

// ---- timing helpers --------------------------------------------------------------------------------------------
static void pass_begin(sphb200_ctx* c) {
    if (!c->timing) return;
    if (!c->ev_created) {
        for (int i = 0; i <= SPH_MAX_PASSES; i++) cudaEventCreate(&c->ev[i]);
        c->ev_created = true;
    }
    c->npass = 0;
    cudaEventRecord(c->ev[0], c->stream);
}
static void pass_mark(sphb200_ctx* c, const char* name) {
    if (!c->timing || c->npass >= SPH_MAX_PASSES) return;
    c->pass_name[c->npass] = name;
    c->npass++;
    cudaEventRecord(c->ev[c->npass], c->stream);
}

// ---- device allocations (guarded in the SPH_DEBUG_BOUNDS build, see ctx.cuh) ----------------------------------------------
#ifdef SPH_DEBUG_BOUNDS
#include <map>
#include <mutex>
namespace {
constexpr size_t DBG_GUARD = 256;
struct DbgAlloc { char* base; size_t bytes; int device; };
std::map<void*, DbgAlloc> g_dbg_allocs;
std::mutex g_dbg_mutex;
}
cudaError_t sph_dev_malloc(void** p, size_t bytes) {
    char* base = nullptr;
    const size_t body = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc((void**)&base, body + 2 * DBG_GUARD);
    if (e != cudaSuccess) return e;
    e = cudaMemset(base, 0xA5, body + 2 * DBG_GUARD);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    *p = base + DBG_GUARD;
    std::lock_guard<std::mutex> lk(g_dbg_mutex);
    g_dbg_allocs[*p] = DbgAlloc{base, bytes, dev};
    return cudaSuccess;
}
cudaError_t sph_dev_free(void* p) {
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lk(g_dbg_mutex);
    auto it = g_dbg_allocs.find(p);
    if (it == g_dbg_allocs.end()) return cudaFree(p);
    char* base = it->second.base;
    g_dbg_allocs.erase(it);
    return cudaFree(base);
}
// guard bytes (the 256 before every allocation and everything from its last requested byte to the end of the rear guard)
// that no longer hold the fill pattern
int sph_debug_guard_errors(long long* bad_bytes, long long* allocations) {
    std::lock_guard<std::mutex> lk(g_dbg_mutex);
    long long bad = 0;
    int cur = 0;
    cudaGetDevice(&cur);
    std::vector<unsigned char> h;
    for (auto& kv : g_dbg_allocs) {
        const DbgAlloc& a = kv.second;
        cudaSetDevice(a.device);
        if (cudaDeviceSynchronize() != cudaSuccess) { cudaSetDevice(cur); return SPH_ERR_CUDA; }
        const size_t body = (a.bytes + 255) & ~(size_t)255;
        h.resize(DBG_GUARD);
        if (cudaMemcpy(h.data(), a.base, DBG_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaSetDevice(cur); return SPH_ERR_CUDA; }
        for (unsigned char b : h) bad += b != 0xA5;
        const size_t tail = body - a.bytes + DBG_GUARD;
        h.resize(tail);
        if (cudaMemcpy(h.data(), a.base + DBG_GUARD + a.bytes, tail, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaSetDevice(cur); return SPH_ERR_CUDA; }
        for (unsigned char b : h) bad += b != 0xA5;
    }
    cudaSetDevice(cur);
    if (bad_bytes) *bad_bytes = bad;
    if (allocations) *allocations = (long long)g_dbg_allocs.size();
    return SPH_OK;
}
#else
cudaError_t sph_dev_malloc(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
cudaError_t sph_dev_free(void* p) { return cudaFree(p); }
int sph_debug_guard_errors(long long* bad_bytes, long long* allocations) {
    if (bad_bytes) *bad_bytes = 0;
    if (allocations) *allocations = -1;   // not a debug build
    return SPH_OK;
}
#endif

template <typename T>
static cudaError_t dalloc(T** p, size_t count) {
    return sph_dev_malloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
}

extern "C" {

#ifdef SPH_DEBUG_BOUNDS
const char* sphb200_version(void) { return "sphb200 0.1 (sm_100a, SPH_DEBUG_BOUNDS)"; }
#else
const char* sphb200_version(void) { return "sphb200 0.1 (sm_100a)"; }
#endif

int sphb200_debug_check_guards(int64_t* bad_bytes, int64_t* allocations) {
    long long b = 0, a = 0;
    int rc = sph_debug_guard_errors(&b, &a);
    if (bad_bytes) *bad_bytes = b;
    if (allocations) *allocations = a;
    return rc;
}

int sphb200_default_params(sph_Params* p) {
    if (!p) return SPH_ERR_INVALID_ARG;
    memset(p, 0, sizeof(*p));
    p->K = 1000.0f;            // PressureFieldSystem.cs:31
    p->G = 1.0f;               // GravityFieldSystem.cs:26
    p->theta = 0.7f;           // GravityFieldSystem.cs:228
    p->target_neighbors = 50;  // ParticleSmoothingSystem.cs:18
    p->max_neighbors = 256;
    p->leaf_max = 4;           // <= 4 bodies per BVH leaf, BoundingVolumeHierarchy.cs:40-83
    p->aabb_mode = 0;
    p->max_grid_bits = 0;
    p->flags = 0;
    return SPH_OK;
}

int sphb200_destroy(sph_handle c) {
    if (!c) return SPH_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    for (int k = 0; k < 2; k++) { sph_dev_free(c->posh[k]); sph_dev_free(c->velm[k]); sph_dev_free(c->orig[k]); sph_dev_free(c->keys[k]); sph_dev_free(c->idx[k]); }
    sph_dev_free(c->posm); sph_dev_free(c->posc); sph_dev_free(c->chunk_counter); sph_dev_free(c->sort_hist); sph_dev_free(c->cell_start); sph_dev_free(c->cell_end); sph_dev_free(c->cell_hmax); sph_dev_free(c->nlist);
    sph_dev_free(c->ncount); sph_dev_free(c->nown); sph_dev_free(c->rho); sph_dev_free(c->press); sph_dev_free(c->cvol); sph_dev_free(c->gradp);
    sph_dev_free(c->grav); sph_dev_free(c->npart); sph_dev_free(c->napprox); sph_dev_free(c->gpart); sph_dev_free(c->tbox); sph_dev_free(c->child); sph_dev_free(c->range);
    sph_dev_free(c->parent); sph_dev_free(c->flag); sph_dev_free(c->mom); sph_dev_free(c->nlo); sph_dev_free(c->nhi); sph_dev_free(c->packed); sph_dev_free(c->bounds);
    sph_dev_free(c->scratch_d); sph_dev_free(c->grid_d); sph_dev_free(c->err_d); sph_dev_free(c->rr_table); sph_dev_free(c->diag_d); sph_dev_free(c->stage_d);
    if (c->err_h) cudaFreeHost(c->err_h);
    if (c->stage_h) cudaFreeHost(c->stage_h);
    if (c->ev_created) for (int i = 0; i <= SPH_MAX_PASSES; i++) cudaEventDestroy(c->ev[i]);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_sph) cudaEventDestroy(c->ev_sph);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return SPH_OK;
}

}  // extern "C"

// Largest grid for n particles: ~2 cells per particle at most, <= 2^8 cells per axis.  A function of the particle count
// (not of the handle's capacity) so that a group rank and a single handle lay the same grid over the same particles.
int sph_grid_bits_for(const sph_Params& p, int64_t n) {
    if (p.max_grid_bits > 0) return p.max_grid_bits;
    int bits = (int)floor(log2(2.0 * (double)std::max<int64_t>(n, 1)) / 3.0);
    return std::min(std::max(bits, 1), 8);
}

int sph_validate_params(const sph_Params& p, int64_t capacity, std::string& err) {
    if (capacity <= 0 || capacity > 0x7fffff00LL / 2) { err = "capacity out of range"; return SPH_ERR_CAPACITY; }
    if (p.max_neighbors <= 0 || p.max_neighbors % 32 != 0 || p.max_neighbors > 1024) { err = "max_neighbors must be a multiple of 32 in [32,1024]"; return SPH_ERR_INVALID_ARG; }
    if (p.leaf_max < 1 || p.leaf_max > 64) { err = "leaf_max must be in [1,64]"; return SPH_ERR_INVALID_ARG; }
    if (p.max_grid_bits < 0 || p.max_grid_bits > 8) { err = "max_grid_bits must be in [0,8]"; return SPH_ERR_INVALID_ARG; }
    return SPH_OK;
}

// Allocate a context: `slots` resident particle slots, neighbor rows / staging for `rows` of them, an LBVH over `tree`
// (global) slots, a cell table for up to `total` particles.  A single handle uses one capacity for all four.
int sph_ctx_create(const sph_Params& p, int device, int64_t slots, int64_t rows, int64_t tree, int64_t total, bool pingpong,
                   sphb200_ctx** out, std::string& err) {
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { err = std::string("no CUDA device (no CPU fallback exists): ") + cudaGetErrorString(e); return SPH_ERR_CUDA; }
    if (device < 0 || device >= ndev) { err = "bad device index"; return SPH_ERR_INVALID_ARG; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) { err = "device is not sm_100-class (library is built for sm_100a only)"; return SPH_ERR_CUDA; }
    sphb200_ctx* c = new (std::nothrow) sphb200_ctx();
    if (!c) return SPH_ERR_CUDA;
    c->p = p; c->device = device; c->cap = slots; c->cap_rows = rows; c->sm_count = prop.multiProcessorCount;
    c->grid_bits_max = sph_grid_bits_for(p, total);
    c->ncell_max = (size_t)1 << (3 * c->grid_bits_max);
    c->gpart_entries = std::max<size_t>(8 * (size_t)rows, std::min<size_t>(32 * (size_t)rows, (size_t)1 << 26));
    size_t cap = (size_t)slots, nr = (size_t)rows, nn = 2 * (size_t)tree;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    {   // the auxiliary stream (LBVH build / gravity-source exchange behind the neighbor pass) gets the highest priority: its
        // blocks are placed first whenever an SM frees resources
        int lo = 0, hi = 0;
        ok = ok && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, hi) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_sph, cudaEventDisableTiming) == cudaSuccess;
    c->stream = c->own_stream;
    for (int k = 0; k < 2 && ok; k++) {
        if (k == 0 || pingpong)
            ok = ok && dalloc(&c->posh[k], cap) == cudaSuccess && dalloc(&c->velm[k], cap) == cudaSuccess && dalloc(&c->orig[k], cap) == cudaSuccess;
        ok = ok && dalloc(&c->keys[k], cap) == cudaSuccess && dalloc(&c->idx[k], cap) == cudaSuccess;
    }
    c->stage_bytes = std::max<size_t>(nr * 14 * 4, (nr + 1) * 8);   // upload: pos3 vel3 mass1 + raw smoothing records (7)
    ok = ok && dalloc(&c->posm, cap) == cudaSuccess && dalloc(&c->posc, cap) == cudaSuccess && dalloc(&c->chunk_counter, 1) == cudaSuccess && dalloc(&c->sort_hist, sph_sort_hist_words(slots)) == cudaSuccess &&
         dalloc(&c->cell_start, c->ncell_max) == cudaSuccess && dalloc(&c->cell_end, c->ncell_max) == cudaSuccess && dalloc(&c->cell_hmax, c->ncell_max) == cudaSuccess &&
         dalloc(&c->nlist, nr * (size_t)p.max_neighbors) == cudaSuccess && dalloc(&c->ncount, cap) == cudaSuccess &&
         dalloc(&c->nown, cap) == cudaSuccess && dalloc(&c->rho, cap) == cudaSuccess && dalloc(&c->press, cap) == cudaSuccess &&
         dalloc(&c->cvol, cap) == cudaSuccess && dalloc(&c->gradp, cap) == cudaSuccess && dalloc(&c->grav, cap) == cudaSuccess &&
         dalloc(&c->npart, cap) == cudaSuccess && dalloc(&c->napprox, cap) == cudaSuccess &&
         dalloc(&c->gpart, c->gpart_entries) == cudaSuccess && dalloc(&c->tbox, 2 * ((size_t)tree / 256 + 2)) == cudaSuccess && dalloc(&c->child, nn) == cudaSuccess &&
         dalloc(&c->range, nn) == cudaSuccess && dalloc(&c->parent, nn) == cudaSuccess && dalloc(&c->flag, nn) == cudaSuccess &&
         dalloc(&c->mom, nn) == cudaSuccess && dalloc(&c->nlo, nn) == cudaSuccess && dalloc(&c->nhi, nn) == cudaSuccess && dalloc(&c->packed, 2 * nn) == cudaSuccess &&
         dalloc(&c->bounds, 16) == cudaSuccess && dalloc(&c->grid_d, 1) == cudaSuccess && dalloc(&c->err_d, ERR_SLOTS) == cudaSuccess &&
         dalloc(&c->rr_table, SPH_RR_TABLE) == cudaSuccess && dalloc(&c->diag_d, 32) == cudaSuccess &&
         sph_dev_malloc(&c->stage_d, c->stage_bytes) == cudaSuccess && cudaMallocHost((void**)&c->err_h, 2 * ERR_SLOTS * sizeof(int32_t)) == cudaSuccess &&
         cudaMallocHost(&c->stage_h, c->stage_bytes) == cudaSuccess;
    if (!ok) {
        err = std::string("allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        sphb200_destroy(c);
        return SPH_ERR_CUDA;
    }
    // radius-ratio table: correctly rounded fp32 pow(target/n, 1/3f) (ParticleSmoothingSystem.cs:49-50); the same
    // expression as the oracle's orc_radius_ratio, so the controller is bit-identical on both sides.
    std::vector<float> rr(SPH_RR_TABLE, 1.0f);
    for (int k = 1; k < SPH_RR_TABLE; k++) {
        float ratio = p.target_neighbors / (float)k;
        rr[k] = (float)pow((double)ratio, (double)(1.0f / 3.0f));
    }
    uint32_t b0[16] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    ok = cudaMemcpy(c->rr_table, rr.data(), SPH_RR_TABLE * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(c->bounds, b0, sizeof(b0), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemset(c->err_d, 0, ERR_SLOTS * sizeof(int32_t)) == cudaSuccess &&
         cudaMemset(c->nown, 0, cap * sizeof(int32_t)) == cudaSuccess && cudaMemset(c->ncount, 0, cap * sizeof(int32_t)) == cudaSuccess &&
         cudaMemset(c->npart, 0, cap * sizeof(int32_t)) == cudaSuccess && cudaMemset(c->napprox, 0, cap * sizeof(int32_t)) == cudaSuccess;
    if (!ok) { err = "initialisation failed"; sphb200_destroy(c); return SPH_ERR_CUDA; }
    *out = c;
    return SPH_OK;
}

extern "C" {

int sphb200_create(const sph_Params* params, int64_t capacity, int device, sph_handle* out) {
    if (!out) return SPH_ERR_INVALID_ARG;
    *out = nullptr;
    sph_Params p;
    if (params) p = *params; else sphb200_default_params(&p);
    int rc = sph_validate_params(p, capacity, g_create_err);
    if (rc) return rc;
    return sph_ctx_create(p, device, capacity, capacity, capacity, capacity, true, out, g_create_err);
}

const char* sphb200_last_error(sph_handle c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int sphb200_set_stream(sph_handle c, void* s) {
    if (!c) return SPH_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->aux_stream);
    cudaStreamSynchronize(c->stream);
    c->tree_join_pending = false;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return SPH_OK;
}

int sphb200_get_stream(sph_handle c, void** s) {
    if (!c || !s) return SPH_ERR_INVALID_ARG;
    *s = (void*)c->stream;
    return SPH_OK;
}

static int check_errflags(sphb200_ctx* c);
static int join_tree(sphb200_ctx* c);

int sphb200_sync(sph_handle c) {
    if (!c) return SPH_ERR_INVALID_ARG;
    SPH_CK(c, cudaSetDevice(c->device));
    return check_errflags(c);  // synchronises the stream and reports sticky asynchronous errors (list overflow, ...)
}

int sphb200_count(sph_handle c, int64_t* n, int64_t* cap) {
    if (!c) return SPH_ERR_INVALID_ARG;
    if (n) *n = c->n;
    if (cap) *cap = c->cap;
    return SPH_OK;
}
int sphb200_get_params(sph_handle c, sph_Params* out) {
    if (!c || !out) return SPH_ERR_INVALID_ARG;
    *out = c->p;
    out->max_grid_bits = c->grid_bits_max;
    return SPH_OK;
}
int sphb200_launch_count(sph_handle c, int64_t* l) {
    if (!c || !l) return SPH_ERR_INVALID_ARG;
    *l = c->launches;
    return SPH_OK;
}
int sphb200_enable_timing(sph_handle c, int en) {
    if (!c) return SPH_ERR_INVALID_ARG;
    c->timing = en != 0;
    c->npass = 0;
    return SPH_OK;
}
int sphb200_get_timings(sph_handle c, const char** names, float* ms, int cap) {
    if (!c) return SPH_ERR_INVALID_ARG;
    if (!c->timing || c->npass == 0) return 0;
    cudaSetDevice(c->device);
    cudaEventSynchronize(c->ev[c->npass]);
    int k = 0;
    for (; k < c->npass && k < cap; k++) {
        if (names) names[k] = c->pass_name[k];
        float t = 0;
        cudaEventElapsedTime(&t, c->ev[k], c->ev[k + 1]);
        if (ms) ms[k] = t;
    }
    return k;
}
int sphb200_fp32_peak(sph_handle c, double* tf) {
    if (!c || !tf) return SPH_ERR_INVALID_ARG;
    SPH_CK(c, cudaSetDevice(c->device));
    return sph_fp32_peak(c, tf);
}

int sphb200_set_target_range(sph_handle c, int64_t t0, int64_t t1) {
    if (!c) return SPH_ERR_INVALID_ARG;
    ARG_CHECK(c, t0 >= 0, "t0 < 0");
    c->t0 = t0;
    c->t1 = t1;
    return SPH_OK;
}

// ---- upload ---------------------------------------------------------------------------------------------------------
}  // extern "C"

// Stage the component arrays of `n` bodies (body indices orig0 .. orig0+n-1) and pack them into the resident SoA at slot 0.
// Asynchronous on the context's stream; the mass range lands in bounds[12..13] (ordered uints) for sph_upload_finish.
int sph_upload_core(sphb200_ctx* c, int64_t n, uint32_t orig0, const void* pos, int pos_stride, const void* vel, int vel_stride,
                    const void* mass, int mass_stride, const void* smoothing, int smoothing_stride) {
    ARG_CHECK(c, n >= 0, "n < 0");
    ARG_CHECK(c, n == 0 || (pos && vel && mass && smoothing), "null component array");
    ARG_CHECK(c, pos_stride >= 12 && vel_stride >= 12 && mass_stride >= 4 && smoothing_stride >= 4, "stride too small");
    ARG_CHECK(c, (size_t)n * 14 * 4 <= c->stage_bytes || n == 0, "upload larger than the staging buffer");
    SPH_CK(c, cudaSetDevice(c->device));
    SPH_CK(c, cudaStreamSynchronize(c->aux_stream));
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    c->tree_join_pending = c->tree_fresh = c->tree_hint = false;
    c->cur = 0;
    c->resident = true; c->lists_valid = c->pressure_valid = c->gravity_valid = c->tree_valid = c->h_updated = false;
    c->sorted_valid = c->lists_fresh = false;
    c->nown_aligned = true;
    c->sph_ready = false;
    // asynchronous error flags are sticky from one upload to the next (results after an overflow are tainted)
    SPH_CK(c, cudaMemsetAsync(c->err_d, 0, ERR_SLOTS * sizeof(int32_t), c->stream));
    const uint32_t mm0[3] = {0xffffffffu, 0u, 0u};
    SPH_CK(c, cudaMemcpyAsync(c->bounds + 12, mm0, sizeof(mm0), cudaMemcpyHostToDevice, c->stream));
    if (n == 0) return SPH_OK;
    // Staging layout (device): pos[3n] vel[3n] mass[n] h[n] nown[n].  Arrays with their natural stride are copied
    // straight from the caller's memory (one DMA when it is pinned); strided arrays are packed through pinned staging.
    float* st = (float*)c->stage_h;
    float* sd = (float*)c->stage_d;
    bool has_nown = smoothing_stride >= (int)sizeof(sph_ParticleSmoothing);
    const char* pp = (const char*)pos; const char* vp = (const char*)vel; const char* mp = (const char*)mass; const char* sp = (const char*)smoothing;
    auto put = [&](size_t off_words, const char* src, int stride, int words) -> cudaError_t {
        size_t bytes = (size_t)n * words * 4;
        if (stride == words * 4) return cudaMemcpyAsync(sd + off_words, src, bytes, cudaMemcpyHostToDevice, c->stream);
        float* dst = st + off_words;
        for (int64_t i = 0; i < n; i++) memcpy(dst + (size_t)words * i, src + (size_t)i * stride, (size_t)words * 4);
        return cudaMemcpyAsync(sd + off_words, dst, bytes, cudaMemcpyHostToDevice, c->stream);
    };
    SPH_CK(c, put(0, pp, pos_stride, 3));
    SPH_CK(c, put(3 * (size_t)n, vp, vel_stride, 3));
    SPH_CK(c, put(6 * (size_t)n, mp, mass_stride, 1));
    int nown_mode = has_nown ? 1 : 0;
    if (smoothing_stride == (int)sizeof(sph_ParticleSmoothing)) {
        // the component array as it is: one DMA, h and the neighbor count are picked out on the device
        SPH_CK(c, cudaMemcpyAsync(sd + 7 * (size_t)n, sp, (size_t)n * sizeof(sph_ParticleSmoothing), cudaMemcpyHostToDevice, c->stream));
        nown_mode = 2;
    } else {
        SPH_CK(c, put(7 * (size_t)n, sp, smoothing_stride, 1));
        if (has_nown) SPH_CK(c, put(8 * (size_t)n, sp + 24, smoothing_stride, 1));
    }
    return sph_launch_pack_upload(c, n, nown_mode, orig0);   // also reduces the mass range into bounds[12..13]
}

// Blocks; equal masses (the reference spawner, ParticleAuthoring.cs:208) select the kernels that hoist the mass multiply.
int sph_upload_finish(sphb200_ctx* c) {
    SPH_CK(c, cudaMemcpyAsync(c->err_h + 4, c->bounds + 12, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    const uint32_t lo = (uint32_t)c->err_h[4], hi = (uint32_t)c->err_h[5];
    c->equal_mass = lo == hi;
    const uint32_t bits = (lo & 0x80000000u) ? (lo & 0x7fffffffu) : ~lo;   // ord2f on the host
    memcpy(&c->common_mass, &bits, 4);
    memcpy(&c->h_bound, &c->err_h[6], 4);   // non-negative floats order like their bit patterns
    return SPH_OK;
}

extern "C" {

int sphb200_upload(sph_handle c, int64_t n, const void* pos, int pos_stride, const void* vel, int vel_stride, const void* mass,
                   int mass_stride, const void* smoothing, int smoothing_stride) {
    if (!c) return SPH_ERR_INVALID_ARG;
    if (n > c->cap) { c->err = "n exceeds capacity"; return SPH_ERR_CAPACITY; }
    int rc = sph_upload_core(c, n, 0u, pos, pos_stride, vel, vel_stride, mass, mass_stride, smoothing, smoothing_stride);
    if (rc) return rc;
    c->n = n;
    // single handle: the resident slots are the global slots
    c->skeys = c->tkeys = c->keys[1];
    c->gsrc = c->posm; c->gsrc_n = n;
    c->tree_n = n; c->tree_g0 = 0; c->tree_g1 = n; c->tree_off = 0; c->row_base = 0;
    c->grid_bits_max = sph_grid_bits_for(c->p, n);   // a function of n, not of the capacity (same grid as any other handle / group)
    return sph_upload_finish(c);
}

// ---- stages ----------------------------------------------------------------------------------------------------------
static int check_errflags(sphb200_ctx* c) {
    SPH_CK(c, cudaMemcpyAsync(c->err_h, c->err_d, ERR_SLOTS * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    if (c->err_h[ERR_TREE_STACK]) { c->err = "LBVH traversal stack overflow"; return SPH_ERR_TREE_STACK; }
    if (c->err_h[ERR_NEIGHBOR_OVERFLOW]) {
        c->err = "neighbor list overflow: a particle has " + std::to_string(c->err_h[ERR_NEIGHBOR_OVERFLOW]) +
                 " neighbors > max_neighbors=" + std::to_string(c->p.max_neighbors) +
                 " (lists truncated: the pressure gradient and the softened near-pair gravity of that particle miss the rest; density and h control stay complete)";
        return SPH_ERR_NEIGHBOR_OVERFLOW;
    }
    return SPH_OK;
}

#define NEED_RESIDENT(c)                                                  \
    do {                                                                  \
        if (!(c)) return SPH_ERR_INVALID_ARG;                             \
        if (!(c)->resident) { (c)->err = "no particles uploaded"; return SPH_ERR_STATE; } \
        SPH_CK(c, cudaSetDevice((c)->device));                            \
    } while (0)

int sphb200_smoothing_update(sph_handle c) {
    NEED_RESIDENT(c);
    if (c->n == 0) return SPH_OK;
    if (!c->nown_aligned) {
        c->err = "smoothing_update after a sort without a neighbor pass: the own-support counts are in the previous slot order "
                 "(call it first in a step, as ParticleSmoothingSystem runs, or after build_neighbors)";
        return SPH_ERR_STATE;
    }
    int rc = join_tree(c);   // a pending LBVH build on the auxiliary stream reads posh.w
    if (rc) return rc;
    rc = sph_launch_smoothing_bounds(c, true);
    if (rc) return rc;
    // h <- 0.5 h (1 + (target/n_own)^(1/3)), n_own >= 1 (ParticleSmoothingSystem.cs:46-59): the largest growth of one update
    c->h_bound *= 0.5f * (1.0f + cbrtf(fmaxf(c->p.target_neighbors, 1.0f))) * 1.001f;
    c->h_updated = true;
    c->sorted_valid = c->lists_fresh = c->pressure_valid = c->gravity_valid = false;
    c->sph_ready = false;
    return SPH_OK;
}

// The auxiliary-stream LBVH build must have finished before the main stream overwrites anything it reads or writes.
static int join_tree(sphb200_ctx* c) {
    if (c->tree_join_pending) {
        SPH_CK(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        c->tree_join_pending = false;
    }
    return SPH_OK;
}

// After a fresh sort: start the LBVH build on the auxiliary stream if the caller announced tree gravity for this step
// (sphb200_prepare_gravity / sphb200_step), so that it overlaps the neighbor pass.
static int maybe_fork_tree(sphb200_ctx* c) {
    if (!c->tree_hint || c->tree_fresh) return SPH_OK;
    SPH_CK(c, cudaEventRecord(c->ev_fork, c->stream));
    SPH_CK(c, cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
    int rc = sph_launch_tree_build(c, c->hint_dt, c->aux_stream);
    if (rc) return rc;
    SPH_CK(c, cudaEventRecord(c->ev_join, c->aux_stream));
    c->tree_fresh = true;
    c->tree_dt = c->hint_dt;
    c->tree_join_pending = true;
    c->tree_hint = false;
    return SPH_OK;
}

static int ensure_sorted(sphb200_ctx* c) {
    if (c->sorted_valid) return SPH_OK;  // sorted for the current positions already
    int rc;
    if ((rc = join_tree(c))) return rc;
    if (!c->h_updated) { rc = sph_launch_smoothing_bounds(c, false); if (rc) return rc; c->h_updated = true; }
    rc = sph_launch_sort_and_cells(c);
    if (rc) return rc;
    c->sorted_valid = true;
    c->lists_valid = c->lists_fresh = c->tree_valid = c->tree_fresh = false;  // slot indices changed
    c->sph_ready = false;
    c->nown_aligned = false;
    return SPH_OK;
}

int sphb200_build_neighbors(sph_handle c) {
    NEED_RESIDENT(c);
    if (c->n == 0) return SPH_OK;
    int rc = ensure_sorted(c);
    if (rc) return rc;
    if ((rc = maybe_fork_tree(c))) return rc;
    rc = sph_launch_neighbors_density(c);
    if (rc) return rc;
    c->lists_valid = c->lists_fresh = c->nown_aligned = true;
    c->pressure_valid = false;
    return SPH_OK;
}

int sphb200_gravity(sph_handle c, int impl, float dt) {
    NEED_RESIDENT(c);
    if (c->n == 0) return SPH_OK;
    c->last_dt = dt;
    if (impl == SPH_GRAVITY_NONE) {
        SPH_CK(c, cudaMemsetAsync(c->grav, 0, (size_t)c->n * sizeof(float4), c->stream));
        SPH_CK(c, cudaMemsetAsync(c->npart, 0, (size_t)c->n * sizeof(int32_t), c->stream));
        SPH_CK(c, cudaMemsetAsync(c->napprox, 0, (size_t)c->n * sizeof(int32_t), c->stream));
        c->gravity_valid = true;
        return SPH_OK;
    }
    if (impl == SPH_GRAVITY_PARTICLE) {
        if (!c->lists_fresh) { c->err = "direct gravity needs this step's neighbor lists (softened near-pair correction): call build_neighbors first"; return SPH_ERR_STATE; }
        int rc = sph_launch_gravity_allpairs(c);
        if (rc) return rc;
        rc = sph_launch_gravity_near(c, SPH_GRAVITY_PARTICLE);
        if (rc) return rc;
        c->gravity_valid = true;
        return SPH_OK;
    }
    if (impl == SPH_GRAVITY_TREE) {
        int rc = ensure_sorted(c);
        if (rc) return rc;
        if (c->tree_fresh && c->tree_dt == dt) {
            if ((rc = join_tree(c))) return rc;       // built on the auxiliary stream during the neighbor pass
        } else {
            if ((rc = join_tree(c))) return rc;
            if ((rc = sph_launch_tree_build(c, dt, c->stream))) return rc;
            c->tree_fresh = true;
            c->tree_dt = dt;
        }
        rc = sph_launch_tree_walk(c);
        if (rc) return rc;
        if (c->p.flags & SPH_FLAG_PM07_SOFTENING) {
            if (!c->lists_fresh) { c->err = "SPH_FLAG_PM07_SOFTENING needs this step's neighbor lists: call build_neighbors first"; return SPH_ERR_STATE; }
            if ((rc = sph_launch_gravity_near(c, SPH_GRAVITY_TREE))) return rc;
        }
        c->gravity_valid = true;
        return SPH_OK;
    }
    c->err = "unknown gravity impl";
    return SPH_ERR_INVALID_ARG;
}

int sphb200_prepare_gravity(sph_handle c, int impl, float dt) {
    NEED_RESIDENT(c);
    c->tree_hint = (impl == SPH_GRAVITY_TREE);
    c->hint_dt = dt;
    return SPH_OK;
}

int sphb200_density(sph_handle c) {
    NEED_RESIDENT(c);
    if (!c->lists_fresh && c->n > 0) { c->err = "density needs build_neighbors in this step (the sum is fused into it)"; return SPH_ERR_STATE; }
    return SPH_OK;
}

int sphb200_pressure(sph_handle c) {
    NEED_RESIDENT(c);
    if (c->n == 0) return SPH_OK;
    if (!c->lists_fresh) { c->err = "pressure needs build_neighbors in this step"; return SPH_ERR_STATE; }
    int rc = sph_launch_pressure(c);
    if (rc) return rc;
    c->pressure_valid = true;
    SPH_CK(c, cudaEventRecord(c->ev_sph, c->stream));
    c->sph_ready = true;
    return SPH_OK;
}

int sphb200_integrate(sph_handle c, float dt) {
    NEED_RESIDENT(c);
    if (c->n == 0) return SPH_OK;
    if (!c->pressure_valid || !c->gravity_valid) { c->err = "integrate needs pressure and gravity of this step"; return SPH_ERR_STATE; }
    int rc = join_tree(c);
    if (rc) return rc;
    rc = sph_launch_integrate(c, dt);
    if (rc) return rc;
    // positions moved: neighbor/tree structures belong to the old positions; they stay downloadable until the next
    // sort but no longer feed any compute stage
    c->h_updated = c->sorted_valid = c->lists_fresh = c->pressure_valid = c->gravity_valid = false;
    return SPH_OK;
}

int sphb200_step(sph_handle c, float dt, int impl) {
    NEED_RESIDENT(c);
    if (c->n == 0) return SPH_OK;
    int rc;
    pass_begin(c);
    if ((rc = sphb200_smoothing_update(c))) return rc;
    pass_mark(c, "smoothing_bounds");
    if (impl == SPH_GRAVITY_TREE) { c->tree_hint = true; c->hint_dt = dt; }
    if ((rc = ensure_sorted(c))) return rc;
    pass_mark(c, "keys_sort_permute_cells");
    if ((rc = maybe_fork_tree(c))) return rc;
    if ((rc = sph_launch_neighbors_density(c))) return rc;
    c->lists_valid = c->lists_fresh = c->nown_aligned = true;
    pass_mark(c, "neighbors_density_eos");
    // the pressure gradient does not depend on gravity: it runs first, so that rho, P and grad P are final early and their
    // downloads (sphb200_download on the auxiliary stream) run beside the gravity pass
    if ((rc = sphb200_pressure(c))) return rc;
    pass_mark(c, "pressure_grad");
    if ((rc = sphb200_gravity(c, impl, dt))) return rc;
    pass_mark(c, impl == SPH_GRAVITY_TREE ? "gravity_tree" : (impl == SPH_GRAVITY_PARTICLE ? "gravity_allpairs" : "gravity_none"));
    if ((rc = sphb200_integrate(c, dt))) return rc;
    pass_mark(c, "integrate");
    return SPH_OK;
}

// ---- download --------------------------------------------------------------------------------------------------------
int sphb200_download(sph_handle c, int field, void* dst, int stride) {
    NEED_RESIDENT(c);
    ARG_CHECK(c, dst || c->n == 0, "null dst");
    if (c->n == 0) return SPH_OK;
    int eb = 0;
    if (field < 0 || field >= SPH_FIELD_COUNT_) { c->err = "unknown field"; return SPH_ERR_INVALID_ARG; }
    // a sph_ParticleSmoothing array with its natural stride is assembled on the device and written by one DMA
    const bool sm_record = field == SPH_FIELD_SMOOTHING && stride == (int)sizeof(sph_ParticleSmoothing);
    // Fields that are final behind the pressure pass (ev_sph) travel on the auxiliary stream: the call returns when ITS copy is
    // done, not when the step is -- a host that asks for them first reads them while the gravity pass is still running.  (h sits
    // in posh.w, which the integrate kernel rewrites with the same value: reading it beside that kernel is harmless.)
    const bool early = c->sph_ready && (field == SPH_FIELD_DENSITY || field == SPH_FIELD_PRESSURE || field == SPH_FIELD_PRESSURE_GRAD ||
                                        field == SPH_FIELD_NEIGHBOR_COUNT || field == SPH_FIELD_SMOOTHING);
    cudaStream_t main_stream = c->stream;
    if (early) {
        SPH_CK(c, cudaStreamWaitEvent(c->aux_stream, c->ev_sph, 0));
        c->stream = c->aux_stream;
    }
    int rc = sph_launch_unpack_field(c, sm_record ? (int)SPH_FIELD_COUNT_ : field, &eb);
    if (rc) { c->stream = main_stream; if (rc == SPH_ERR_INVALID_ARG) c->err = "unknown field"; return rc; }
    int64_t n = c->n;
    const bool direct = stride == eb && (field != SPH_FIELD_SMOOTHING || sm_record);
    cudaError_t ce = cudaMemcpyAsync(direct ? dst : c->stage_h, c->stage_d, (size_t)n * eb, cudaMemcpyDeviceToHost, c->stream);
    if (ce != cudaSuccess) { c->stream = main_stream; c->err = cudaGetErrorString(ce); return SPH_ERR_CUDA; }
    rc = check_errflags(c);  // also synchronises (the early path: the auxiliary stream only; the tree flag is reported by a later call)
    c->stream = main_stream;
    if (direct) return rc;   // natural stride: the DMA wrote the caller's array
    const char* src = (const char*)c->stage_h;
    char* d = (char*)dst;
    switch (field) {
        case SPH_FIELD_SMOOTHING: {
            ARG_CHECK(c, stride >= 4, "stride too small");
            bool full = stride >= (int)sizeof(sph_ParticleSmoothing);
            for (int64_t i = 0; i < n; i++) {
                float h; int32_t no;
                memcpy(&h, src + 8 * i, 4); memcpy(&no, src + 8 * i + 4, 4);
                if (full) {
                    sph_ParticleSmoothing s;
                    s.influenceArea = h; s.supportDomain = 2.0f * h;  // ParticleSmoothing.cs:16-23
                    s.sphereColliderPosRadius[0] = s.sphereColliderPosRadius[1] = s.sphereColliderPosRadius[2] = 0.f;
                    s.sphereColliderPosRadius[3] = 2.0f * h;
                    s.neighbors = no;
                    memcpy(d + (size_t)i * stride, &s, sizeof(s));
                } else memcpy(d + (size_t)i * stride, &h, 4);
            }
            break;
        }
        case SPH_FIELD_GRAVITY: {
            ARG_CHECK(c, stride >= 16, "stride too small");
            int w = stride >= 24 ? 24 : 16;
            for (int64_t i = 0; i < n; i++) memcpy(d + (size_t)i * stride, src + 24 * i, w);
            break;
        }
        default: {
            ARG_CHECK(c, stride >= eb, "stride too small");
            if (stride == eb) memcpy(d, src, (size_t)n * eb);
            else for (int64_t i = 0; i < n; i++) memcpy(d + (size_t)i * stride, src + (size_t)i * eb, eb);
        }
    }
    return rc;
}

}  // extern "C"

// grow-only scratch (the previous contents are not kept)
static int scratch_reserve(sphb200_ctx* c, size_t bytes) {
    if (bytes <= c->scratch_bytes) return SPH_OK;
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    sph_dev_free(c->scratch_d); c->scratch_d = nullptr; c->scratch_bytes = 0;
    bytes += bytes / 8;
    if (sph_dev_malloc(&c->scratch_d, bytes) != cudaSuccess) { cudaGetLastError(); c->err = "allocation failed (download scratch)"; return SPH_ERR_CUDA; }
    c->scratch_bytes = bytes;
    return SPH_OK;
}

extern "C" {

int sphb200_download_neighbors(sph_handle c, int64_t* offsets, int32_t* nbr, int64_t cap, int64_t* total) {
    NEED_RESIDENT(c);
    ARG_CHECK(c, offsets && total, "null argument");
    if (!c->lists_valid && c->n > 0) { c->err = "no neighbor lists"; return SPH_ERR_STATE; }
    int64_t n = c->n;
    offsets[0] = 0;
    if (n == 0) { *total = 0; return SPH_OK; }
    int eb;
    int rc = sph_launch_unpack_field(c, SPH_FIELD_NEIGHBOR_COUNT, &eb);
    if (rc) return rc;
    SPH_CK(c, cudaMemcpyAsync(c->stage_h, c->stage_d, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    const int32_t* cnt = (const int32_t*)c->stage_h;
    int kmax = c->p.max_neighbors;
    for (int64_t i = 0; i < n; i++) offsets[i + 1] = offsets[i] + std::min(cnt[i], kmax);
    *total = offsets[n];
    if (*total > cap || !nbr) return check_errflags(c);
    memcpy(c->stage_h, offsets, (size_t)(n + 1) * 8);
    SPH_CK(c, cudaMemcpyAsync(c->stage_d, c->stage_h, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if ((rc = scratch_reserve(c, (size_t)std::max<int64_t>(*total, 1) * 4))) return rc;
    int32_t* rows_d = (int32_t*)c->scratch_d;
    rc = sph_launch_neighbor_rows_sorted(c, rows_d);
    if (rc == SPH_OK) {
        cudaError_t e = cudaMemcpyAsync(nbr, rows_d, (size_t)*total * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = SPH_ERR_CUDA; }
    }
    if (rc) return rc;
    return check_errflags(c);
}

int sphb200_download_interactions(sph_handle c, const int64_t* offsets, const int32_t* nbr, sph_ParticleInteraction* out) {
    NEED_RESIDENT(c);
    ARG_CHECK(c, offsets, "null argument");
    const int64_t n = c->n;
    if (n == 0) return SPH_OK;
    // the records are W / grad W at the positions the lists were built for (KernelSystem.cs:290-334): after integrate or a
    // smoothing update they would describe other positions than the lists do
    if (!c->lists_fresh) { c->err = "interaction records need this step's neighbor lists (call build_neighbors; integrate invalidates them)"; return SPH_ERR_STATE; }
    ARG_CHECK(c, offsets[0] == 0, "offsets[0] != 0");
    for (int64_t i = 0; i < n; i++) ARG_CHECK(c, offsets[i + 1] >= offsets[i], "offsets not ascending");
    const int64_t total = offsets[n];
    if (total == 0) return SPH_OK;
    ARG_CHECK(c, nbr && out, "null argument");
    for (int64_t e = 0; e < total; e++) ARG_CHECK(c, nbr[e] >= 0 && nbr[e] < n, "neighbor index out of range");
    const size_t off_b = ((size_t)(n + 1) * 8 + 255) & ~(size_t)255, nbr_b = ((size_t)total * 4 + 255) & ~(size_t)255;
    int rc = scratch_reserve(c, off_b + nbr_b + (size_t)total * sizeof(sph_ParticleInteraction));
    if (rc) return rc;
    int64_t* off_d = (int64_t*)c->scratch_d;
    int32_t* nbr_d = (int32_t*)((char*)c->scratch_d + off_b);
    sph_ParticleInteraction* out_d = (sph_ParticleInteraction*)((char*)c->scratch_d + off_b + nbr_b);
    SPH_CK(c, cudaMemcpyAsync(off_d, offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    SPH_CK(c, cudaMemcpyAsync(nbr_d, nbr, (size_t)total * 4, cudaMemcpyHostToDevice, c->stream));
    if ((rc = sph_launch_interactions(c, total, off_d, nbr_d, out_d))) return rc;
    SPH_CK(c, cudaMemcpyAsync(out, out_d, (size_t)total * sizeof(sph_ParticleInteraction), cudaMemcpyDeviceToHost, c->stream));
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    return SPH_OK;
}

int sphb200_download_sort(sph_handle c, uint32_t* order, uint32_t* keys, sph_GridParams* grid) {
    NEED_RESIDENT(c);
    int64_t n = c->n;
    if (order && n) SPH_CK(c, cudaMemcpyAsync(order, c->orig[c->cur], (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (keys && n) SPH_CK(c, cudaMemcpyAsync(keys, c->keys[1], (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (grid) SPH_CK(c, cudaMemcpyAsync(grid, c->grid_d, sizeof(sph_GridParams), cudaMemcpyDeviceToHost, c->stream));
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    return SPH_OK;
}

int sphb200_download_tree(sph_handle c, int32_t* child, int32_t* range, float* moment, float* lo, float* hi) {
    NEED_RESIDENT(c);
    if (!c->tree_valid) { c->err = "no tree: call gravity(SPH_GRAVITY_TREE) first"; return SPH_ERR_STATE; }
    { int rcj = join_tree(c); if (rcj) return rcj; }
    int64_t nn = 2 * c->n - 1;
    if (nn <= 0) return SPH_OK;
    if (child) SPH_CK(c, cudaMemcpyAsync(child, c->child, (size_t)nn * 8, cudaMemcpyDeviceToHost, c->stream));
    if (range) SPH_CK(c, cudaMemcpyAsync(range, c->range, (size_t)nn * 8, cudaMemcpyDeviceToHost, c->stream));
    if (moment) SPH_CK(c, cudaMemcpyAsync(moment, c->mom, (size_t)nn * 16, cudaMemcpyDeviceToHost, c->stream));
    std::vector<float4> tmp;
    if (lo || hi) {
        tmp.resize((size_t)nn);
        for (int pass = 0; pass < 2; pass++) {
            float* dstp = pass == 0 ? lo : hi;
            if (!dstp) continue;
            SPH_CK(c, cudaMemcpyAsync(tmp.data(), pass == 0 ? c->nlo : c->nhi, (size_t)nn * 16, cudaMemcpyDeviceToHost, c->stream));
            SPH_CK(c, cudaStreamSynchronize(c->stream));
            for (int64_t k = 0; k < nn; k++) { dstp[3 * k] = tmp[k].x; dstp[3 * k + 1] = tmp[k].y; dstp[3 * k + 2] = tmp[k].z; }
        }
    }
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    return SPH_OK;
}

int sphb200_diagnostics(sph_handle c, double* out12) {
    NEED_RESIDENT(c);
    ARG_CHECK(c, out12, "null out");
    if (c->n == 0) { for (int k = 0; k < 12; k++) out12[k] = 0; return SPH_OK; }
    return sph_launch_diagnostics(c, out12);
}

int sphb200_field_stats(sph_handle c, double* out12) {
    NEED_RESIDENT(c);
    ARG_CHECK(c, out12, "null out");
    for (int k = 0; k < 12; k++) out12[k] = 0;
    if (c->n == 0) return SPH_OK;
    if (!c->lists_valid) { c->err = "no fields yet: run a step (or build_neighbors) first"; return SPH_ERR_STATE; }
    int rc = sph_launch_field_stats_range(c, 0, (int)c->n);
    if (rc) return rc;
    double tmp[8];
    SPH_CK(c, cudaMemcpyAsync(tmp, c->diag_d + 16, sizeof(tmp), cudaMemcpyDeviceToHost, c->stream));
    SPH_CK(c, cudaStreamSynchronize(c->stream));
    sph_finish_field_stats(tmp, (const uint32_t*)(tmp + 4), c->n, out12);
    return SPH_OK;
}

// ---- snapshots: the host-visible state in body order, one binary file ------------------------------------------------
// header (64 bytes): magic "SPHB200S", u32 version, u32 reserved, i64 n_total, i64 body0, i64 count, i64 steps, pad;
// then pos[3 count] f32, vel[3 count] f32, mass[count] f32, h[count] f32, n_own[count] i32.
static const char kSnapMagic[8] = {'S', 'P', 'H', 'B', '2', '0', '0', 'S'};
struct SnapHeader { char magic[8]; uint32_t version, reserved; int64_t n_total, body0, count, steps, pad[2]; };
static_assert(sizeof(SnapHeader) == 64, "snapshot header is 64 bytes");

extern "C++" int sph_snapshot_write(const char* path, int64_t n_total, int64_t body0, int64_t count, int64_t steps, const float* pos, const float* vel,
                                    const float* mass, const float* h, const int32_t* nown, std::string& err) {
    FILE* f = fopen(path, "wb");
    if (!f) { err = std::string("cannot open ") + path + " for writing"; return SPH_ERR_INVALID_ARG; }
    SnapHeader hd{};
    memcpy(hd.magic, kSnapMagic, 8); hd.version = 1; hd.n_total = n_total; hd.body0 = body0; hd.count = count; hd.steps = steps;
    const size_t c = (size_t)count;
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
    ok = ok && (c == 0 || (fwrite(pos, 12, c, f) == c && fwrite(vel, 12, c, f) == c && fwrite(mass, 4, c, f) == c && fwrite(h, 4, c, f) == c &&
                           fwrite(nown, 4, c, f) == c));
    ok = (fclose(f) == 0) && ok;
    if (!ok) { err = std::string("short write to ") + path; return SPH_ERR_INVALID_ARG; }
    return SPH_OK;
}

// reads the bodies [want0, want0 + want) of a snapshot that covers them
extern "C++" int sph_snapshot_read(const char* path, int64_t want0, int64_t want, int64_t* n_total, std::vector<float>& pos, std::vector<float>& vel,
                                   std::vector<float>& mass, std::vector<sph_ParticleSmoothing>& sm, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return SPH_ERR_INVALID_ARG; }
    SnapHeader hd{};
    bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && memcmp(hd.magic, kSnapMagic, 8) == 0 && hd.version == 1 && hd.count >= 0 && hd.body0 >= 0;
    if (!ok) { fclose(f); err = std::string(path) + " is not a sphb200 snapshot (version 1)"; return SPH_ERR_INVALID_ARG; }
    if (want < 0) { want0 = hd.body0; want = hd.count; }
    if (want0 < hd.body0 || want0 + want > hd.body0 + hd.count) { fclose(f); err = std::string(path) + " does not hold the requested bodies"; return SPH_ERR_INVALID_ARG; }
    *n_total = hd.n_total;
    const size_t c = (size_t)want, skip = (size_t)(want0 - hd.body0), all = (size_t)hd.count;
    pos.resize(3 * c); vel.resize(3 * c); mass.resize(c); sm.assign(c, sph_ParticleSmoothing{});
    std::vector<float> hh(c); std::vector<int32_t> no(c);
    auto rd = [&](void* dst, size_t elem, size_t base_elems) {
        return fseek(f, (long)(sizeof(hd) + base_elems + skip * elem), SEEK_SET) == 0 && (c == 0 || fread(dst, elem, c, f) == c);
    };
    ok = rd(pos.data(), 12, 0) && rd(vel.data(), 12, all * 12) && rd(mass.data(), 4, all * 24) && rd(hh.data(), 4, all * 28) && rd(no.data(), 4, all * 32);
    fclose(f);
    if (!ok) { err = std::string("short read from ") + path; return SPH_ERR_INVALID_ARG; }
    for (size_t i = 0; i < c; i++) {
        sm[i].influenceArea = hh[i]; sm[i].supportDomain = 2.0f * hh[i];
        sm[i].sphereColliderPosRadius[3] = 2.0f * hh[i]; sm[i].neighbors = no[i];
    }
    return SPH_OK;
}

int sphb200_snapshot_save(sph_handle c, const char* path) {
    NEED_RESIDENT(c);
    ARG_CHECK(c, path, "null path");
    const size_t n = (size_t)c->n;
    std::vector<float> pos(3 * n), vel(3 * n), mass(n), sm2(2 * n), h(n);
    std::vector<int32_t> no(n);
    int rc = SPH_OK;
    if (n) {
        if ((rc = sphb200_download(c, SPH_FIELD_TRANSLATION, pos.data(), 12)) && rc != SPH_ERR_NEIGHBOR_OVERFLOW) return rc;
        if ((rc = sphb200_download(c, SPH_FIELD_VELOCITY, vel.data(), 12)) && rc != SPH_ERR_NEIGHBOR_OVERFLOW) return rc;
        if ((rc = sphb200_download(c, SPH_FIELD_MASS, mass.data(), 4)) && rc != SPH_ERR_NEIGHBOR_OVERFLOW) return rc;
        std::vector<sph_ParticleSmoothing> sm(n);
        if ((rc = sphb200_download(c, SPH_FIELD_SMOOTHING, sm.data(), (int)sizeof(sph_ParticleSmoothing))) && rc != SPH_ERR_NEIGHBOR_OVERFLOW) return rc;
        for (size_t i = 0; i < n; i++) { h[i] = sm[i].influenceArea; no[i] = sm[i].neighbors; }
    }
    return sph_snapshot_write(path, c->n, 0, c->n, 0, pos.data(), vel.data(), mass.data(), h.data(), no.data(), c->err);
}

int sphb200_snapshot_load(sph_handle c, const char* path) {
    if (!c) return SPH_ERR_INVALID_ARG;
    ARG_CHECK(c, path, "null path");
    std::vector<float> pos, vel, mass; std::vector<sph_ParticleSmoothing> sm;
    int64_t n_total = 0;
    int rc = sph_snapshot_read(path, 0, -1, &n_total, pos, vel, mass, sm, c->err);
    if (rc) return rc;
    if ((int64_t)mass.size() != n_total) { c->err = "the snapshot holds a slice of a group; load it with sphb200_group_snapshot_load"; return SPH_ERR_INVALID_ARG; }
    return sphb200_upload(c, n_total, pos.data(), 12, vel.data(), 12, mass.data(), 4, sm.data(), (int)sizeof(sph_ParticleSmoothing));
}

int sphb200_device_ptr(sph_handle c, const char* name, void** ptr, int64_t* bytes) {
    if (!c || !name || !ptr) return SPH_ERR_INVALID_ARG;
    struct { const char* nm; void* p; size_t el; } tab[] = {
        {"posh", c->posh[c->cur], 16}, {"velm", c->velm[c->cur], 16}, {"posm", c->posm, 16}, {"rho", c->rho, 4},
        {"press", c->press, 4}, {"cvol", c->cvol, 4}, {"gradp", c->gradp, 16}, {"grav", c->grav, 16}, {"nown", c->nown, 4},
        {"orig", c->orig[c->cur], 4}, {"ncount", c->ncount, 4}, {"npart", c->npart, 4}, {"napprox", c->napprox, 4},
    };
    for (auto& t : tab)
        if (strcmp(t.nm, name) == 0) { *ptr = t.p; if (t.p == (void*)c->posh[c->cur]) c->h_bound = INFINITY;   /* the caller may write h */ if (bytes) *bytes = (int64_t)(t.el * (size_t)c->cap); return SPH_OK; }
    c->err = std::string("unknown array name: ") + name;
    return SPH_ERR_INVALID_ARG;
}

}  // extern "C"
