// group.cu -- multi-GPU: Morton-range domain decomposition with halo exchange behind the C ABI (sphb200_group_*).
//
// Replaces: nothing the reference has (its parallelism is shared-memory job threads, UP/Collision/World/Broadphase.cs:163, and
// it is capped at 2^24-2 bodies, UP/Dynamics/Simulation/Scheduler.cs:24-41); this is how the replacement scales N.
//
// A group is `world` ranks, one per GPU.  Either ONE process drives all of them (sphb200_group_create: a C# host calls it
// once; transport = NCCL communicators from ncclCommInitAll, or -- when several ranks share a device, as the single-GPU
// tests do -- in-process peer copies), or every process drives ONE rank (sphb200_group_create_rank: torchrun; transport =
// NCCL with a unique id the host distributes).  libnccl is resolved at run time (dlopen): the library loads without it.
//
// Per step every rank keeps only its Morton range plus a halo (kernels_group.cu has the device side):
//   1. h update + bounds on the own particles; all-reduce of the bounds (integer min/max/sum: order-free) -> one grid
//   2. keys; histogram over 2^18 key-prefix bins, all-reduced; work-balanced splitters; particles that left the rank's key
//      range migrate (all-to-all of 48-byte records over NVLink; a handful per step once the run is settled)
//   3. local stable radix sort of the received set (source-rank order + stable sort = the global stable order)
//   3b. the GRAVITY LANE starts here on the auxiliary stream / communicator and runs behind passes 4-6: gravity sources and keys of
//      the own range straight from the sorted records, all-gathered in global sorted order; tree gravity builds the LBVH nodes of
//      the own range, all-gathers the few records that describe the nodes straddling rank boundaries (k_top_tree finishes them
//      identically on every rank) and exchanges a LOCALLY ESSENTIAL TREE: the sender tests every finished node's parent against
//      the box of the positions each other rank walks and ships only the records a walk of that rank can reach (k_let_*,
//      kernels_group.cu: ~8 % of the nodes at 8 ranks x 2 M) -- all-gathering the node array would move 64 B x N per step
//   4. halo: own particles whose cell stencil touches another rank's cells go to that rank (32-byte records);
//      extended set [low halo | own | high halo] = a sorted subsequence of the global order, with its own cell table
//   5. neighbor rows + density for the own targets; (m/rho)P of the halo particles follows through the same lists
//   6. pressure gradient; then the gravity lane is joined and the own targets walk the tree (or run the all-pairs kernel)
//   7. integration of the own particles
// Three host synchronisations per step (migration counts, halo counts and -- on the gravity lane only, while the device runs the
// neighbor pass -- the tree-node counts): NCCL needs the message sizes on the host.
// Results are bit-identical to the single-GPU step for tree gravity (tests/test_group.py), <= 1e-6 for all-pairs (the
// source-split partial sums depend on the number of targets per rank).
#include "group.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

// ---- NCCL, resolved at run time ------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
    void* lib = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;   // optional (NCCL >= 2.18)
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    // a libnccl the process already holds (torch's bundled one under torchrun) wins, then the system library
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (api.lib) break; }
    if (!api.lib) for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) { api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return &api; }
#define NSYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name); if (!api.field) { api.err = std::string("missing symbol ") + name; api.lib = nullptr; return &api; }
    NSYM(GetUniqueId, "ncclGetUniqueId") NSYM(CommInitRank, "ncclCommInitRank") NSYM(CommInitAll, "ncclCommInitAll")
    NSYM(CommDestroy, "ncclCommDestroy") NSYM(AllReduce, "ncclAllReduce") NSYM(AllGather, "ncclAllGather") NSYM(Send, "ncclSend")
    NSYM(Recv, "ncclRecv") NSYM(GroupStart, "ncclGroupStart") NSYM(GroupEnd, "ncclGroupEnd") NSYM(GetErrorString, "ncclGetErrorString")
#undef NSYM
    *(void**)(&api.CommSplit) = dlsym(api.lib, "ncclCommSplit");
    return &api;
}

struct GroupRank {
    sphb200_ctx* c = nullptr;
    int rank = 0, device = 0;
    ncclComm_t comm = nullptr;       // the communicator the collectives currently use (comm_main or comm_aux)
    ncclComm_t comm_main = nullptr, comm_aux = nullptr;   // comm_aux: a split of comm_main for the auxiliary stream (null: no overlap)
    cudaStream_t main_stream = nullptr;
    // device buffers
    uint4 *mig_send = nullptr, *mig_recv = nullptr;   // 6 x 16 B per own slot: migration records (48 B) / result records (96 B)
    uint8_t* dest = nullptr;
    uint32_t *perm = nullptr, *tot256 = nullptr;
    uint32_t *hmask = nullptr, *hlist = nullptr, *hcnt = nullptr, *htot = nullptr;
    uint4 *halo_send = nullptr, *halo_recv = nullptr;
    float* cv_send = nullptr;
    uint32_t* hist = nullptr;
    unsigned long long* bpart = nullptr;   // run sums of the histograms (k_bin_partials)
    uint32_t* bmask = nullptr;             // owners around every bin (k_bin_owner_mask)
    int64_t* split_d = nullptr;      // g0[0..world], sbin[0..world]
    uint32_t* cnt_d = nullptr;       // [world][world] count matrices (migration, halo, download)
    float4* posm_g = nullptr;
    uint32_t* keys_g = nullptr;
    float4* bnd = nullptr;           // [world][2][SPH_TOP_LEAF][2]
    // locally essential tree exchange (kernels_group.cu k_let_*)
    uint32_t *let_mask = nullptr, *let_cnt = nullptr, *let_cursor = nullptr, *let_soff = nullptr, *let_box = nullptr, *lcnt_d = nullptr;
    float4 *let_send = nullptr, *let_recv = nullptr;
    float4 *tposh = nullptr, *tvelm = nullptr;   // own particles in sorted order (what the LBVH build reads)
    uint32_t* lcnt_h = nullptr;      // pinned [world][world]
    void* red_scratch = nullptr;     // in-process transport
    size_t red_bytes = 0;
    // pinned host mirrors
    uint32_t* cnt_h = nullptr;
    int64_t* split_h = nullptr;
    sph_GridParams* grid_h = nullptr;
    double* diag_h = nullptr;
    // host state
    int64_t n_own = 0, own0 = 0, n_ext = 0, n_hsend = 0;
    int64_t body0 = 0, nbody = 0, n_res = 0;
    cudaEvent_t ev = nullptr;
};

}  // namespace

struct sphb200_group {
    sph_Params p{};
    int world = 1, nlocal = 1, rank0 = 0;
    bool local = false;               // in-process transport (every rank is local)
    int64_t cap_total = 0, cap_own = 0, cap_halo = 0, cap_ext = 0, cap_let = 0;
    int64_t last_let = 0;             // walk-node records received by the local ranks in the last tree step (-1: full all-gather)
    int64_t n_total = 0, chunk = 0;
    std::vector<GroupRank> r;
    std::vector<int64_t> g0;          // global slot ranges of the ranks, world+1 (after the first step)
    bool resident = false, stepped = false, res_valid = false;
    int64_t steps = 0;
    int64_t last_migrated = 0, last_halo = 0;
    // timing (events on the first local rank's stream)
    bool timing = false;
    int npass = 0;
    const char* pass_name[SPH_MAX_PASSES];
    cudaEvent_t tev[SPH_MAX_PASSES + 1];
    bool tev_created = false;
    // the auxiliary lane of the step (gravity sources / tree exchange): events on the first local rank's auxiliary stream
    int naux = 0, aux_fork_pass = 0;
    const char* aux_name[12];
    cudaEvent_t aev[13];
    std::string err;
};

static std::string g_group_err;

namespace {

#define G_FAIL(g, code, msg) do { (g)->err = (msg); return (code); } while (0)
#define G_CUDA(g, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { (g)->err = std::string(#call) + ": " + cudaGetErrorString(e__); return SPH_ERR_CUDA; } } while (0)
#define G_NCCL(g, call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) { (g)->err = std::string(#call) + ": " + nccl_api()->GetErrorString(r__); return SPH_ERR_NCCL; } } while (0)
// a launcher of a rank's context failed: lift its message
#define G_RC(g, R, call) do { int rc__ = (call); if (rc__) { (g)->err = (R).c->err; return rc__; } } while (0)
#define FOR_RANKS(g, R) for (GroupRank& R : (g)->r)

// every local stream waits for everything issued so far on every other local stream
int barrier_local(sphb200_group* g) {
    if (g->nlocal <= 1) return SPH_OK;
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_CUDA(g, cudaEventRecord(R.ev, R.c->stream)); }
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        FOR_RANKS(g, Q) if (&Q != &R) G_CUDA(g, cudaStreamWaitEvent(R.c->stream, Q.ev, 0));
    }
    return SPH_OK;
}

ncclDataType_t nc_type(int t) { return t == GR_U32 ? ncclUint32 : t == GR_U64 ? ncclUint64 : ncclFloat64; }
ncclRedOp_t nc_op(int o) { return o == GR_SUM ? ncclSum : o == GR_MIN ? ncclMin : ncclMax; }
size_t gr_size(int t) { return t == GR_U32 ? 4 : 8; }

// ---- collectives over all ranks; every function takes one pointer per LOCAL rank (in the order of g->r) -----------------
int g_allreduce(sphb200_group* g, void* const* bufs, size_t count, int dtype, int op) {
    if (g->world == 1 || count == 0) return SPH_OK;
    if (!g->local) {
        NcclApi* N = nccl_api();
        G_NCCL(g, N->GroupStart());
        for (int l = 0; l < g->nlocal; l++) {
            GroupRank& R = g->r[l];
            G_CUDA(g, cudaSetDevice(R.device));
            G_NCCL(g, N->AllReduce(bufs[l], bufs[l], count, nc_type(dtype), nc_op(op), R.comm, R.c->stream));
        }
        G_NCCL(g, N->GroupEnd());
        return SPH_OK;
    }
    const size_t bytes = count * gr_size(dtype);
    int rc;
    if ((rc = barrier_local(g))) return rc;
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        if (bytes * g->world > R.red_bytes) G_FAIL(g, SPH_ERR_INVALID_ARG, "all-reduce larger than the scratch buffer");
        G_CUDA(g, cudaSetDevice(R.device));
        for (int q = 0; q < g->world; q++)
            G_CUDA(g, cudaMemcpyAsync((char*)R.red_scratch + (size_t)q * bytes, bufs[q], bytes, cudaMemcpyDefault, R.c->stream));
    }
    if ((rc = barrier_local(g))) return rc;
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        G_CUDA(g, cudaSetDevice(R.device));
        G_RC(g, R, grk_reduce_ranks(R.c, R.red_scratch, g->world, count, dtype, op, bufs[l]));
    }
    return SPH_OK;
}

// in place: rank q's segment of `bytes` bytes sits at bufs[l] + q*bytes
int g_allgather(sphb200_group* g, void* const* bufs, size_t bytes) {
    if (g->world == 1 || bytes == 0) return SPH_OK;
    if (!g->local) {
        NcclApi* N = nccl_api();
        G_NCCL(g, N->GroupStart());
        for (int l = 0; l < g->nlocal; l++) {
            GroupRank& R = g->r[l];
            G_CUDA(g, cudaSetDevice(R.device));
            G_NCCL(g, N->AllGather((char*)bufs[l] + (size_t)R.rank * bytes, bufs[l], bytes, ncclChar, R.comm, R.c->stream));
        }
        G_NCCL(g, N->GroupEnd());
        return SPH_OK;
    }
    int rc;
    if ((rc = barrier_local(g))) return rc;
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        G_CUDA(g, cudaSetDevice(R.device));
        for (int q = 0; q < g->world; q++)
            if (q != l) G_CUDA(g, cudaMemcpyAsync((char*)bufs[l] + (size_t)q * bytes, (char*)bufs[q] + (size_t)q * bytes, bytes, cudaMemcpyDefault, R.c->stream));
    }
    return barrier_local(g);
}

// in place, unequal segments: rank q's segment is elements [off[q], off[q+1]) of every rank's array
int g_allgatherv(sphb200_group* g, void* const* bufs, const int64_t* off, size_t elem) {
    if (g->world == 1) return SPH_OK;
    if (!g->local) {
        NcclApi* N = nccl_api();
        G_NCCL(g, N->GroupStart());
        for (int l = 0; l < g->nlocal; l++) {
            GroupRank& R = g->r[l];
            G_CUDA(g, cudaSetDevice(R.device));
            const int me = R.rank;
            for (int q = 0; q < g->world; q++) {
                if (q == me) continue;
                if (off[me + 1] > off[me]) G_NCCL(g, N->Send((char*)bufs[l] + (size_t)off[me] * elem, (size_t)(off[me + 1] - off[me]) * elem, ncclChar, q, R.comm, R.c->stream));
                if (off[q + 1] > off[q]) G_NCCL(g, N->Recv((char*)bufs[l] + (size_t)off[q] * elem, (size_t)(off[q + 1] - off[q]) * elem, ncclChar, q, R.comm, R.c->stream));
            }
        }
        G_NCCL(g, N->GroupEnd());
        return SPH_OK;
    }
    int rc;
    if ((rc = barrier_local(g))) return rc;
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        G_CUDA(g, cudaSetDevice(R.device));
        for (int q = 0; q < g->world; q++)
            if (q != l && off[q + 1] > off[q])
                G_CUDA(g, cudaMemcpyAsync((char*)bufs[l] + (size_t)off[q] * elem, (char*)bufs[q] + (size_t)off[q] * elem, (size_t)(off[q + 1] - off[q]) * elem, cudaMemcpyDefault, R.c->stream));
    }
    return barrier_local(g);
}

// M[p*world+q] = elements p sends to q.  Rank me sends send[l] + soff(me,q) and receives from p at recv[l] + roff[l][p].
// soff(me,q) = sum_{q'<q} M[me][q'].
int g_alltoallv(sphb200_group* g, const void* const* send, void* const* recv, const std::vector<std::vector<int64_t>>& roff,
                const uint32_t* M, size_t elem) {
    const int W = g->world;
    auto soff = [&](int me, int q) { int64_t s = 0; for (int k = 0; k < q; k++) s += M[me * W + k]; return s; };
    if (!g->local) {
        NcclApi* N = nccl_api();
        for (int l = 0; l < g->nlocal; l++) {      // self part: a device copy
            GroupRank& R = g->r[l];
            const int me = R.rank;
            G_CUDA(g, cudaSetDevice(R.device));
            if (M[me * W + me])
                G_CUDA(g, cudaMemcpyAsync((char*)recv[l] + (size_t)roff[l][me] * elem, (const char*)send[l] + (size_t)soff(me, me) * elem,
                                          (size_t)M[me * W + me] * elem, cudaMemcpyDeviceToDevice, R.c->stream));
        }
        if (W == 1) return SPH_OK;
        G_NCCL(g, N->GroupStart());
        for (int l = 0; l < g->nlocal; l++) {
            GroupRank& R = g->r[l];
            const int me = R.rank;
            G_CUDA(g, cudaSetDevice(R.device));
            for (int q = 0; q < W; q++) {
                if (q == me) continue;
                if (M[me * W + q]) G_NCCL(g, N->Send((const char*)send[l] + (size_t)soff(me, q) * elem, (size_t)M[me * W + q] * elem, ncclChar, q, R.comm, R.c->stream));
                if (M[q * W + me]) G_NCCL(g, N->Recv((char*)recv[l] + (size_t)roff[l][q] * elem, (size_t)M[q * W + me] * elem, ncclChar, q, R.comm, R.c->stream));
            }
        }
        G_NCCL(g, N->GroupEnd());
        return SPH_OK;
    }
    int rc;
    if ((rc = barrier_local(g))) return rc;
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        G_CUDA(g, cudaSetDevice(R.device));
        for (int p = 0; p < W; p++)
            if (M[p * W + l])
                G_CUDA(g, cudaMemcpyAsync((char*)recv[l] + (size_t)roff[l][p] * elem, (const char*)send[p] + (size_t)soff(p, l) * elem,
                                          (size_t)M[p * W + l] * elem, cudaMemcpyDefault, R.c->stream));
    }
    return barrier_local(g);
}

int sync_all(sphb200_group* g) {
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_CUDA(g, cudaStreamSynchronize(R.c->stream)); }
    return SPH_OK;
}

// Auxiliary lane: between aux_begin and aux_end everything a rank issues (kernels of its context, collectives) goes to its
// auxiliary stream / communicator and runs concurrently with what is issued on the main stream afterwards, until aux_join.
bool aux_available(sphb200_group* g) {
    if (g->world == 1) return false;
    if (g->local) return getenv("SPHB200_GROUP_NO_OVERLAP") == nullptr;
    for (auto& R : g->r) if (!R.comm_aux) return false;
    return true;
}
int aux_begin(sphb200_group* g) {
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_CUDA(g, cudaEventRecord(c->ev_fork, c->stream));
        G_CUDA(g, cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        R.main_stream = c->stream;
        c->stream = c->aux_stream;
        if (!g->local) R.comm = R.comm_aux;
    }
    return SPH_OK;
}
// leave the auxiliary lane without closing it (no join event): the lane is resumed later in the step
int aux_pause(sphb200_group* g) {
    FOR_RANKS(g, R) { R.c->stream = R.main_stream; R.comm = R.comm_main; }
    return SPH_OK;
}
int aux_resume(sphb200_group* g) {
    FOR_RANKS(g, R) {
        R.main_stream = R.c->stream;
        R.c->stream = R.c->aux_stream;
        if (!g->local) R.comm = R.comm_aux;
    }
    return SPH_OK;
}
int aux_end(sphb200_group* g) {
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_CUDA(g, cudaEventRecord(c->ev_join, c->stream));
        c->stream = R.main_stream;
        R.comm = R.comm_main;
    }
    return SPH_OK;
}
int aux_join(sphb200_group* g) {
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_CUDA(g, cudaStreamWaitEvent(R.c->stream, R.c->ev_join, 0)); }
    return SPH_OK;
}

template <typename T>
std::vector<void*> ptrs(sphb200_group* g, T f) {
    std::vector<void*> v;
    FOR_RANKS(g, R) v.push_back((void*)f(R));
    return v;
}

void pass_begin(sphb200_group* g) {
    if (!g->timing) return;
    GroupRank& R = g->r[0];
    cudaSetDevice(R.device);
    if (!g->tev_created) {
        for (int i = 0; i <= SPH_MAX_PASSES; i++) cudaEventCreate(&g->tev[i]);
        for (int i = 0; i < 13; i++) cudaEventCreate(&g->aev[i]);
        g->tev_created = true;
    }
    g->npass = 0;
    g->naux = 0;
    cudaEventRecord(g->tev[0], R.c->stream);
}
// a mark on whatever stream rank 0 currently issues to (the auxiliary lane between aux_begin/aux_resume and aux_pause/aux_end)
void aux_mark(sphb200_group* g, const char* name) {
    if (!g->timing || g->naux >= 12) return;
    GroupRank& R = g->r[0];
    cudaSetDevice(R.device);
    if (g->naux == 0) g->aux_fork_pass = g->npass;   // lane time is counted from the last main-lane mark (the fork point)
    g->aux_name[g->naux++] = name;
    cudaEventRecord(g->aev[g->naux], R.c->stream);
}
void pass_mark(sphb200_group* g, const char* name) {
    if (!g->timing || g->npass >= SPH_MAX_PASSES) return;
    GroupRank& R = g->r[0];
    cudaSetDevice(R.device);
    g->pass_name[g->npass++] = name;
    cudaEventRecord(g->tev[g->npass], R.c->stream);
}

template <typename T> cudaError_t dal(T** p, size_t count) { return sph_dev_malloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)); }

void free_rank(GroupRank& R) {
    if (R.c) cudaSetDevice(R.device);
    sph_dev_free(R.mig_send); sph_dev_free(R.mig_recv); sph_dev_free(R.dest); sph_dev_free(R.perm); sph_dev_free(R.tot256); sph_dev_free(R.hmask); sph_dev_free(R.hlist);
    sph_dev_free(R.hcnt); sph_dev_free(R.htot); sph_dev_free(R.halo_send); sph_dev_free(R.halo_recv); sph_dev_free(R.cv_send); sph_dev_free(R.hist); sph_dev_free(R.bpart); sph_dev_free(R.bmask);
    sph_dev_free(R.tposh); sph_dev_free(R.tvelm); sph_dev_free(R.let_mask); sph_dev_free(R.let_cnt); sph_dev_free(R.let_cursor); sph_dev_free(R.let_soff); sph_dev_free(R.let_box); sph_dev_free(R.lcnt_d); sph_dev_free(R.let_send); sph_dev_free(R.let_recv);
    if (R.lcnt_h) cudaFreeHost(R.lcnt_h);
    sph_dev_free(R.split_d); sph_dev_free(R.cnt_d); sph_dev_free(R.posm_g); sph_dev_free(R.keys_g); sph_dev_free(R.bnd); sph_dev_free(R.red_scratch);
    if (R.c) { sph_dev_free(R.c->top_nodes); sph_dev_free(R.c->front_nodes); sph_dev_free(R.c->top_counts); R.c->top_nodes = nullptr; R.c->front_nodes = nullptr; R.c->top_counts = nullptr; }
    if (R.cnt_h) cudaFreeHost(R.cnt_h);
    if (R.split_h) cudaFreeHost(R.split_h);
    if (R.grid_h) cudaFreeHost(R.grid_h);
    if (R.diag_h) cudaFreeHost(R.diag_h);
    if (R.ev) cudaEventDestroy(R.ev);
    if (R.comm_aux && nccl_api()->lib) nccl_api()->CommDestroy(R.comm_aux);
    if (R.comm_main && nccl_api()->lib) nccl_api()->CommDestroy(R.comm_main);
    if (R.c) sphb200_destroy(R.c);
    R = GroupRank();
}

int alloc_rank(sphb200_group* g, GroupRank& R) {
    const int W = g->world;
    std::string err;
    int rc = sph_ctx_create(g->p, R.device, g->cap_ext, g->cap_own, g->cap_total, g->cap_total, false, &R.c, err);
    if (rc) { g->err = err; return rc; }
    const size_t own = (size_t)g->cap_own, halo = (size_t)(2 * g->cap_halo), N = (size_t)g->cap_total;
    bool ok = cudaSetDevice(R.device) == cudaSuccess;
    R.red_bytes = (size_t)W * std::max<size_t>(2 * (size_t)SPH_NBINS * 4, 4096);
    ok = ok && dal(&R.mig_send, 6 * own) == cudaSuccess && dal(&R.mig_recv, 6 * own) == cudaSuccess && dal(&R.dest, own) == cudaSuccess &&
         dal(&R.perm, own) == cudaSuccess && dal(&R.tot256, 256) == cudaSuccess && dal(&R.hmask, own) == cudaSuccess &&
         dal(&R.hlist, halo) == cudaSuccess && dal(&R.hcnt, grk_halo_cnt_words(g->cap_own)) == cudaSuccess && dal(&R.htot, 256) == cudaSuccess &&
         dal(&R.halo_send, 2 * halo) == cudaSuccess && dal(&R.halo_recv, 2 * halo) == cudaSuccess && dal(&R.cv_send, halo) == cudaSuccess &&
         dal(&R.hist, 2 * (size_t)SPH_NBINS) == cudaSuccess && dal(&R.bpart, 2 * (size_t)SPH_NBINS / 1024) == cudaSuccess && dal(&R.bmask, (size_t)SPH_NBINS) == cudaSuccess && dal(&R.split_d, 2 * (size_t)(W + 1)) == cudaSuccess &&
         dal(&R.cnt_d, (size_t)W * W) == cudaSuccess && dal(&R.posm_g, N) == cudaSuccess && dal(&R.keys_g, N) == cudaSuccess &&
         dal(&R.bnd, (size_t)W * 2 * SPH_TOP_LEAF * 2) == cudaSuccess &&
         dal(&R.tposh, own) == cudaSuccess && dal(&R.tvelm, own) == cudaSuccess && dal(&R.let_mask, 2 * own) == cudaSuccess && dal(&R.let_cnt, (size_t)W) == cudaSuccess && dal(&R.let_cursor, (size_t)W) == cudaSuccess &&
         dal(&R.let_soff, (size_t)W) == cudaSuccess && dal(&R.let_box, (size_t)W * 8) == cudaSuccess && dal(&R.lcnt_d, (size_t)W * W) == cudaSuccess &&
         (W == 1 || (dal(&R.let_send, 3 * (size_t)g->cap_let) == cudaSuccess && dal(&R.let_recv, 3 * (size_t)g->cap_let) == cudaSuccess)) &&
         cudaMallocHost((void**)&R.lcnt_h, (size_t)W * W * sizeof(uint32_t)) == cudaSuccess &&
         (!g->local || sph_dev_malloc(&R.red_scratch, R.red_bytes) == cudaSuccess) &&
         dal(&R.c->top_nodes, (size_t)W * SPH_TOP_CAP) == cudaSuccess && dal(&R.c->front_nodes, (size_t)W * SPH_TOP_CAP) == cudaSuccess &&
         dal(&R.c->top_counts, (size_t)W * 4) == cudaSuccess &&
         cudaMallocHost((void**)&R.cnt_h, (size_t)W * W * sizeof(uint32_t)) == cudaSuccess &&
         cudaMallocHost((void**)&R.split_h, 2 * (size_t)(W + 1) * sizeof(int64_t)) == cudaSuccess &&
         cudaMallocHost((void**)&R.grid_h, sizeof(sph_GridParams)) == cudaSuccess &&
         cudaMallocHost((void**)&R.diag_h, 32 * sizeof(double)) == cudaSuccess &&
         cudaEventCreateWithFlags(&R.ev, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { g->err = std::string("group allocation failed: ") + cudaGetErrorString(cudaGetLastError()); return SPH_ERR_CUDA; }
    R.c->top_rank = R.rank;
    return SPH_OK;
}

int create_common(const sph_Params* params, int64_t capacity, int world, int rank0, int nlocal, const int* devices, bool local,
                  sphb200_group** out) {
    sph_Params p;
    if (params) p = *params; else sphb200_default_params(&p);
    int rc = sph_validate_params(p, capacity, g_group_err);
    if (rc) return rc;
    if (world < 1 || world > SPH_MAX_RANKS) { g_group_err = "world size must be in [1,32]"; return SPH_ERR_INVALID_ARG; }
    sphb200_group* g = new (std::nothrow) sphb200_group();
    if (!g) return SPH_ERR_CUDA;
    g->p = p; g->world = world; g->nlocal = nlocal; g->rank0 = rank0; g->local = local;
    g->cap_total = capacity;
    // Own slots: twice the even share (the splitters balance work, not counts), at least a few thousand; halo slots on either
    // side: half of that plus 64 k -- a small group of a small test may need every other particle as halo.
    const int64_t share = (capacity + world - 1) / world;
    g->cap_own = world == 1 ? capacity : std::min<int64_t>(capacity, 2 * share + 4096);
    g->cap_halo = world == 1 ? 0 : std::min<int64_t>(capacity, g->cap_own / 2 + 65536);
    g->cap_ext = g->cap_own + 2 * g->cap_halo;
    g->cap_let = 2 * g->cap_own;      // walk-node records a rank may send / receive per step (beyond it: the full all-gather)
    g->r.resize(nlocal);
    for (int l = 0; l < nlocal; l++) { g->r[l].rank = rank0 + l; g->r[l].device = devices[l]; }
    for (int l = 0; l < nlocal; l++)
        if ((rc = alloc_rank(g, g->r[l]))) { g_group_err = g->err; for (auto& R : g->r) free_rank(R); delete g; return rc; }
    g->g0.assign(world + 1, 0);
    *out = g;
    return SPH_OK;
}

}  // namespace

extern "C" {

int sphb200_group_unique_id(void* id128) {
    if (!id128) return SPH_ERR_INVALID_ARG;
    NcclApi* N = nccl_api();
    if (!N->lib) { g_group_err = N->err; return SPH_ERR_NCCL; }
    ncclUniqueId id;
    if (N->GetUniqueId(&id) != ncclSuccess) { g_group_err = "ncclGetUniqueId failed"; return SPH_ERR_NCCL; }
    memcpy(id128, &id, sizeof(id));
    return SPH_OK;
}

int sphb200_group_create(const sph_Params* params, int64_t capacity, int ndev, const int* devices, sph_group* out) {
    if (!out || ndev < 1 || ndev > SPH_MAX_RANKS) { g_group_err = "bad argument"; return SPH_ERR_INVALID_ARG; }
    *out = nullptr;
    std::vector<int> devs(ndev);
    bool dup = false;
    for (int l = 0; l < ndev; l++) {
        devs[l] = devices ? devices[l] : l;
        for (int k = 0; k < l; k++) dup = dup || devs[k] == devs[l];
    }
    const char* env = getenv("SPHB200_GROUP_TRANSPORT");
    // several ranks on one device (tests on a single GPU): NCCL refuses duplicate devices -> in-process peer copies
    const bool local = ndev > 1 && (dup || (env && strcmp(env, "local") == 0));
    sphb200_group* g = nullptr;
    int rc = create_common(params, capacity, ndev, 0, ndev, devs.data(), local, &g);
    if (rc) return rc;
    if (ndev > 1 && !local) {
        NcclApi* N = nccl_api();
        if (!N->lib) { g_group_err = N->err; sphb200_group_destroy(g); return SPH_ERR_NCCL; }
        std::vector<ncclComm_t> comms(ndev);
        ncclResult_t r = N->CommInitAll(comms.data(), ndev, devs.data());
        if (r != ncclSuccess) { g_group_err = std::string("ncclCommInitAll: ") + N->GetErrorString(r); sphb200_group_destroy(g); return SPH_ERR_NCCL; }
        for (int l = 0; l < ndev; l++) g->r[l].comm = g->r[l].comm_main = comms[l];
        if (N->CommSplit && !getenv("SPHB200_GROUP_NO_OVERLAP")) {
            bool okk = N->GroupStart() == ncclSuccess;
            for (int l = 0; l < ndev && okk; l++) { cudaSetDevice(devs[l]); okk = N->CommSplit(comms[l], 0, l, &g->r[l].comm_aux, nullptr) == ncclSuccess; }
            okk = (N->GroupEnd() == ncclSuccess) && okk;
            if (!okk) for (int l = 0; l < ndev; l++) g->r[l].comm_aux = nullptr;
        }
    }
    *out = g;
    return SPH_OK;
}

int sphb200_group_create_rank(const sph_Params* params, int64_t capacity, const void* id128, int world, int rank, int device, sph_group* out) {
    if (!out || rank < 0 || rank >= world || (world > 1 && !id128)) { g_group_err = "bad argument"; return SPH_ERR_INVALID_ARG; }
    *out = nullptr;
    sphb200_group* g = nullptr;
    int rc = create_common(params, capacity, world, rank, 1, &device, false, &g);
    if (rc) return rc;
    if (world > 1) {
        NcclApi* N = nccl_api();
        if (!N->lib) { g_group_err = N->err; sphb200_group_destroy(g); return SPH_ERR_NCCL; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        cudaSetDevice(device);
        ncclResult_t r = N->CommInitRank(&g->r[0].comm, world, id, rank);
        if (r != ncclSuccess) { g_group_err = std::string("ncclCommInitRank: ") + N->GetErrorString(r); sphb200_group_destroy(g); return SPH_ERR_NCCL; }
        g->r[0].comm_main = g->r[0].comm;
        // a second communicator over the same ranks for the auxiliary stream (collective call: every rank makes it or none --
        // the environment switch must be set for all ranks alike)
        if (N->CommSplit && !getenv("SPHB200_GROUP_NO_OVERLAP") && N->CommSplit(g->r[0].comm, 0, rank, &g->r[0].comm_aux, nullptr) != ncclSuccess)
            g->r[0].comm_aux = nullptr;
    }
    *out = g;
    return SPH_OK;
}

int sphb200_group_destroy(sph_group g) {
    if (!g) return SPH_ERR_INVALID_ARG;
    for (auto& R : g->r) if (R.c) { cudaSetDevice(R.device); cudaStreamSynchronize(R.c->stream); }
    if (g->tev_created) { cudaSetDevice(g->r[0].device); for (int i = 0; i <= SPH_MAX_PASSES; i++) cudaEventDestroy(g->tev[i]); }
    for (auto& R : g->r) free_rank(R);
    delete g;
    return SPH_OK;
}

const char* sphb200_group_last_error(sph_group g) { return g ? g->err.c_str() : g_group_err.c_str(); }

int sphb200_group_body_range(sph_group g, int64_t n_total, int64_t* body0, int64_t* count) {
    if (!g || n_total < 0) return SPH_ERR_INVALID_ARG;
    const int64_t chunk = (n_total + g->world - 1) / g->world;
    const int64_t b0 = std::min<int64_t>((int64_t)g->rank0 * chunk, n_total);
    const int64_t b1 = std::min<int64_t>((int64_t)(g->rank0 + g->nlocal) * chunk, n_total);
    if (body0) *body0 = b0;
    if (count) *count = b1 - b0;
    return SPH_OK;
}

// ---- upload -----------------------------------------------------------------------------------------------------------
int sphb200_group_upload(sph_group g, int64_t n_total, const void* pos, int pos_stride, const void* vel, int vel_stride,
                         const void* mass, int mass_stride, const void* smoothing, int smoothing_stride) {
    if (!g) return SPH_ERR_INVALID_ARG;
    if (n_total < 0) G_FAIL(g, SPH_ERR_INVALID_ARG, "n < 0");
    if (n_total > g->cap_total) G_FAIL(g, SPH_ERR_CAPACITY, "n exceeds capacity");
    g->n_total = n_total;
    g->chunk = std::max<int64_t>((n_total + g->world - 1) / g->world, 1);
    int64_t pb0 = 0;
    sphb200_group_body_range(g, n_total, &pb0, nullptr);
    FOR_RANKS(g, R) {
        R.body0 = std::min<int64_t>((int64_t)R.rank * g->chunk, n_total);
        R.nbody = std::min<int64_t>((int64_t)(R.rank + 1) * g->chunk, n_total) - R.body0;
        const int64_t o = R.body0 - pb0;
        auto at = [&](const void* p, int stride) { return p ? (const void*)((const char*)p + (size_t)o * stride) : p; };
        int rc = sph_upload_core(R.c, R.nbody, (uint32_t)R.body0, at(pos, pos_stride), pos_stride, at(vel, vel_stride), vel_stride,
                                 at(mass, mass_stride), mass_stride, at(smoothing, smoothing_stride), smoothing_stride);
        if (rc) { g->err = R.c->err; return rc; }
        sphb200_ctx* c = R.c;
        R.n_own = R.nbody; R.own0 = 0; R.n_ext = R.nbody; R.n_hsend = 0; R.n_res = 0;
        c->n = R.nbody; c->t0 = 0; c->t1 = R.nbody; c->row_base = 0;
        c->grid_bits_max = sph_grid_bits_for(c->p, n_total);
        // weights of the first balance: no neighbor / interaction counts yet
        G_CUDA(g, cudaMemsetAsync(c->ncount, 0, (size_t)c->cap * sizeof(int32_t), c->stream));
        G_CUDA(g, cudaMemsetAsync(c->npart, 0, (size_t)c->cap * sizeof(int32_t), c->stream));
        G_CUDA(g, cudaMemsetAsync(c->napprox, 0, (size_t)c->cap * sizeof(int32_t), c->stream));
    }
    // one mass range for the whole group
    int rc;
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->bounds + 12; }); if ((rc = g_allreduce(g, b.data(), 1, GR_U32, GR_MIN))) return rc; }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->bounds + 13; }); if ((rc = g_allreduce(g, b.data(), 1, GR_U32, GR_MAX))) return rc; }
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_RC(g, R, sph_upload_finish(R.c)); }
    if (n_total == 0) FOR_RANKS(g, R) R.c->equal_mass = false;
    g->resident = true; g->stepped = false; g->res_valid = false;
    for (int r = 0; r <= g->world; r++) g->g0[r] = std::min<int64_t>((int64_t)r * g->chunk, n_total);
    return SPH_OK;
}

// ---- the step -----------------------------------------------------------------------------------------------------------
int sphb200_group_step(sph_group g, float dt, int impl) {
    if (!g) return SPH_ERR_INVALID_ARG;
    if (!g->resident) G_FAIL(g, SPH_ERR_STATE, "no particles uploaded");
    if (impl != SPH_GRAVITY_TREE && impl != SPH_GRAVITY_PARTICLE && impl != SPH_GRAVITY_NONE) G_FAIL(g, SPH_ERR_INVALID_ARG, "unknown gravity impl");
    if (g->n_total == 0) return SPH_OK;
    const int W = g->world;
    const int64_t N = g->n_total;
    int rc;
    g->res_valid = false;
    pass_begin(g);

    // 1. smoothing-length update + bounds of the own particles; one grid for the whole group
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_RC(g, R, sph_launch_bounds_range(c, c->posh[0] + R.own0, c->nown + R.own0, (int)R.n_own, true));
    }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->bounds; }); if ((rc = g_allreduce(g, b.data(), 3, GR_U32, GR_MIN))) return rc; }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->bounds + 3; }); if ((rc = g_allreduce(g, b.data(), 4, GR_U32, GR_MAX))) return rc; }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->bounds + 8; }); if ((rc = g_allreduce(g, b.data(), 1, GR_U64, GR_SUM))) return rc; }
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_RC(g, R, sph_launch_grid_setup(R.c, N)); }
    pass_mark(g, "smoothing_bounds");

    // 2. keys, ownership, migration
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_RC(g, R, sph_launch_keys(c, c->posh[0] + R.own0, (int)R.n_own, c->keys[1]));
        G_RC(g, R, grk_bin_hist(c, c->keys[1], c->ncount + R.own0, c->npart + R.own0, c->napprox + R.own0, (int)R.n_own, R.hist));
    }
    pass_mark(g, "keys_hist");
    { auto b = ptrs(g, [](GroupRank& R) { return R.hist; }); if ((rc = g_allreduce(g, b.data(), 2 * (size_t)SPH_NBINS, GR_U32, GR_SUM))) return rc; }
    pass_mark(g, "allreduce_hist");
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_RC(g, R, grk_splitters(c, R.hist, W, R.split_d, R.bpart, R.bmask));
        G_RC(g, R, grk_dest(c, c->keys[1], (int)R.n_own, R.split_d, W, R.dest));
        G_CUDA(g, cudaMemsetAsync(R.tot256, 0, 256 * sizeof(uint32_t), c->stream));
        G_RC(g, R, sph_launch_digit_pass(c, nullptr, nullptr, R.dest, (int)R.n_own, 0, nullptr, R.perm, R.tot256, c->stream));
        G_CUDA(g, cudaMemcpyAsync(R.cnt_d + (size_t)R.rank * W, R.tot256, W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    }
    { auto b = ptrs(g, [](GroupRank& R) { return R.cnt_d; }); if ((rc = g_allgather(g, b.data(), W * sizeof(uint32_t)))) return rc; }
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        G_CUDA(g, cudaMemcpyAsync(R.cnt_h, R.cnt_d, (size_t)W * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, R.c->stream));
        G_CUDA(g, cudaMemcpyAsync(R.split_h, R.split_d, 2 * (size_t)(W + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, R.c->stream));
        G_CUDA(g, cudaMemcpyAsync(R.grid_h, R.c->grid_d, sizeof(sph_GridParams), cudaMemcpyDeviceToHost, R.c->stream));
    }
    pass_mark(g, "splitters_partition");
    if ((rc = sync_all(g))) return rc;                                   // ---- host sync 1: migration counts
    const uint32_t* M = g->r[0].cnt_h;
    const int64_t* sp = g->r[0].split_h;
    for (int r = 0; r <= W; r++) g->g0[r] = sp[r];
    std::vector<std::vector<int64_t>> roff(g->nlocal, std::vector<int64_t>(W + 1, 0));
    g->last_migrated = 0;
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        const int me = R.rank;
        for (int p = 0; p < W; p++) { roff[l][p + 1] = roff[l][p] + M[p * W + me]; if (p != me) g->last_migrated += M[p * W + me]; }
        const int64_t n_new = roff[l][W];
        if (n_new != g->g0[me + 1] - g->g0[me]) G_FAIL(g, SPH_ERR_STATE, "internal: migration counts disagree with the splitters");
        if (n_new > g->cap_own) G_FAIL(g, SPH_ERR_CAPACITY, "a rank would own " + std::to_string(n_new) + " particles > its capacity " + std::to_string(g->cap_own));
    }
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_RC(g, R, grk_mig_pack(c, c->posh[0] + R.own0, c->velm[0] + R.own0, c->orig[0] + R.own0, c->nown + R.own0, c->keys[1], R.perm,
                                (int)R.n_own, R.mig_send));
    }
    {
        auto s = ptrs(g, [](GroupRank& R) { return R.mig_send; });
        auto d = ptrs(g, [](GroupRank& R) { return R.mig_recv; });
        if ((rc = g_alltoallv(g, s.data(), d.data(), roff, M, 48))) return rc;
    }
    pass_mark(g, "migrate");
    // 6a. gravity sources: all-gather of (x,y,z,m) in global sorted order; tree gravity also all-gathers the keys, builds the LBVH
    // nodes of the own range, all-gathers the packed nodes and finishes the nodes that straddle rank boundaries.  None of it
    // depends on the neighbor pass: it is issued first, on the auxiliary stream and communicator, and runs behind pass 5.
    const char* gname = impl == SPH_GRAVITY_TREE ? "gravity_tree" : impl == SPH_GRAVITY_PARTICLE ? "gravity_allpairs" : "gravity_none";
    // Part A (before the neighbor pass is issued): source / key all-gathers, LBVH build of the own range, top lists, and -- the
    // walk nodes being far too many to all-gather (64 B x N per step) -- the selection of the nodes every other rank can reach
    // (locally essential tree, kernels_group.cu k_let_*) with its count matrix on the way to the host.
    const bool let_on = impl == SPH_GRAVITY_TREE && W > 1 && getenv("SPHB200_GROUP_NO_LET") == nullptr;
    const GroupRank* l0 = &g->r[0];
    auto gravity_part_a = [&]() -> int {
        FOR_RANKS(g, R) {
            G_CUDA(g, cudaSetDevice(R.device));
            sphb200_ctx* c = R.c;
            // from the sorted migration records: the extended resident set (posm, keys[0]) does not exist yet
            G_RC(g, R, grk_sorted_sources(c, R.mig_recv, c->idx[1], c->keys[1], (int)R.n_own, R.posm_g + g->g0[R.rank],
                                          impl == SPH_GRAVITY_TREE ? R.keys_g + g->g0[R.rank] : nullptr, R.tposh, R.tvelm));
            c->gsrc = R.posm_g; c->gsrc_n = N;
        }
        { auto b = ptrs(g, [](GroupRank& R) { return R.posm_g; }); if ((rc = g_allgatherv(g, b.data(), g->g0.data(), 16))) return rc; }
        if (impl != SPH_GRAVITY_TREE) { aux_mark(g, "gravity_aux_sources"); return SPH_OK; }
        { auto b = ptrs(g, [](GroupRank& R) { return R.keys_g; }); if ((rc = g_allgatherv(g, b.data(), g->g0.data(), 4))) return rc; }
        FOR_RANKS(g, R) {
            G_CUDA(g, cudaSetDevice(R.device));
            sphb200_ctx* c = R.c;
            c->tkeys = R.keys_g; c->tree_n = N; c->tree_g0 = g->g0[R.rank]; c->tree_g1 = g->g0[R.rank + 1];
            c->tree_posh = R.tposh; c->tree_velm = R.tvelm; c->tree_src_off = -g->g0[R.rank];
            if (W > 1) G_CUDA(g, cudaMemsetAsync(c->top_counts + 4 * R.rank, 0, 4 * sizeof(int32_t), c->stream));
            if (l0 == &R) aux_mark(g, "gravity_aux_sources_keys");
            G_RC(g, R, sph_launch_tree_build(c, dt, c->stream));
            if (W > 1) G_RC(g, R, grk_boundary(c, R.tposh, R.tvelm, (int)R.n_own, R.bnd + (size_t)R.rank * 2 * SPH_TOP_LEAF * 2));
            if (l0 == &R) aux_mark(g, "gravity_aux_lbvh_build");
        }
        if (W > 1) {
            { auto b = ptrs(g, [](GroupRank& R) { return R.c->top_nodes; }); if ((rc = g_allgather(g, b.data(), SPH_TOP_CAP * sizeof(TopNode)))) return rc; }
            { auto b = ptrs(g, [](GroupRank& R) { return R.c->front_nodes; }); if ((rc = g_allgather(g, b.data(), SPH_TOP_CAP * sizeof(FrontNode)))) return rc; }
            { auto b = ptrs(g, [](GroupRank& R) { return R.c->top_counts; }); if ((rc = g_allgather(g, b.data(), 4 * sizeof(int32_t)))) return rc; }
            { auto b = ptrs(g, [](GroupRank& R) { return R.bnd; }); if ((rc = g_allgather(g, b.data(), 2 * SPH_TOP_LEAF * 2 * sizeof(float4)))) return rc; }
            aux_mark(g, "gravity_aux_top_lists");
        }
        if (let_on) {
            FOR_RANKS(g, R) {
                G_CUDA(g, cudaSetDevice(R.device));
                // the slots this rank walks: its targets and the companions that complete its first and last 32-slot group
                const int64_t lo = g->g0[R.rank] & ~(int64_t)31, hi = std::min<int64_t>(N, (g->g0[R.rank + 1] + 31) & ~(int64_t)31);
                G_RC(g, R, grk_let_box(R.c, R.posm_g, (int)lo, R.n_own > 0 ? (int)hi : (int)lo, R.let_box + 8 * (size_t)R.rank));
            }
            { auto b = ptrs(g, [](GroupRank& R) { return R.let_box; }); if ((rc = g_allgather(g, b.data(), 8 * sizeof(uint32_t)))) return rc; }
            FOR_RANKS(g, R) {
                G_CUDA(g, cudaSetDevice(R.device));
                G_RC(g, R, grk_let_mask(R.c, R.let_box, W, R.rank, R.let_mask, R.let_cnt));
                G_CUDA(g, cudaMemcpyAsync(R.lcnt_d + (size_t)R.rank * W, R.let_cnt, W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, R.c->stream));
            }
            { auto b = ptrs(g, [](GroupRank& R) { return R.lcnt_d; }); if ((rc = g_allgather(g, b.data(), W * sizeof(uint32_t)))) return rc; }
            FOR_RANKS(g, R) {
                G_CUDA(g, cudaSetDevice(R.device));
                G_CUDA(g, cudaMemcpyAsync(R.lcnt_h, R.lcnt_d, (size_t)W * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, R.c->stream));
            }
            aux_mark(g, "gravity_aux_let_select");
        }
        return SPH_OK;
    };
    // Part B (issued once the host has the count matrix; the device is busy with the neighbor pass meanwhile): the selected walk
    // nodes travel to the ranks that can reach them and the straddling nodes are finished.
    auto gravity_part_b = [&]() -> int {
        if (impl != SPH_GRAVITY_TREE || W == 1) return SPH_OK;
        bool let = let_on;
        std::vector<std::vector<int64_t>> loff(g->nlocal, std::vector<int64_t>(W + 1, 0));
        const uint32_t* L = g->r[0].lcnt_h;
        if (let) {
            FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_CUDA(g, cudaStreamSynchronize(R.c->stream)); }   // ---- host sync 3 (this lane only)
            // every rank takes the same decision from the same matrix
            for (int p = 0; p < W && let; p++) {
                int64_t out = 0, in = 0;
                for (int q = 0; q < W; q++) { out += L[p * W + q]; in += L[q * W + p]; }
                if (out > g->cap_let || in > g->cap_let) let = false;
            }
        }
        if (let) {
            g->last_let = 0;
            for (int l = 0; l < g->nlocal; l++) {
                GroupRank& R = g->r[l];
                const int me = R.rank;
                uint32_t soff[SPH_MAX_RANKS];
                uint32_t acc = 0;
                for (int q = 0; q < W; q++) { soff[q] = acc; acc += L[me * W + q]; }
                for (int p = 0; p < W; p++) loff[l][p + 1] = loff[l][p] + L[p * W + me];
                g->last_let += loff[l][W];
                G_CUDA(g, cudaSetDevice(R.device));
                G_CUDA(g, cudaMemcpyAsync(R.let_soff, soff, W * sizeof(uint32_t), cudaMemcpyHostToDevice, R.c->stream));
                G_RC(g, R, grk_let_pack(R.c, R.let_mask, W, R.let_soff, R.let_cursor, R.let_send));
            }
            aux_mark(g, "gravity_aux_hostsync_pack");
            {
                auto sd = ptrs(g, [](GroupRank& R) { return R.let_send; });
                auto rv = ptrs(g, [](GroupRank& R) { return R.let_recv; });
                if ((rc = g_alltoallv(g, sd.data(), rv.data(), loff, L, 48))) return rc;
            }
            aux_mark(g, "gravity_aux_alltoall");
            for (int l = 0; l < g->nlocal; l++) {
                GroupRank& R = g->r[l];
                G_CUDA(g, cudaSetDevice(R.device));
                if (getenv("SPHB200_GROUP_LET_POISON")) {
                    // test aid: every node record this rank neither built nor received reads as NaN / child -1, so a walk that
                    // stepped outside its locally essential tree could not go unnoticed
                    const int64_t a0 = g->g0[R.rank], a1 = g->g0[R.rank + 1];
                    const int64_t seg[3][2] = {{0, a0}, {std::min<int64_t>(a1, N - 1), N - 1 + a0}, {N - 1 + a1, 2 * N - 1}};
                    for (auto& sg : seg)
                        if (sg[1] > sg[0]) G_CUDA(g, cudaMemsetAsync(R.c->packed + 2 * sg[0], 0xff, (size_t)(sg[1] - sg[0]) * 2 * sizeof(float4), R.c->stream));
                }
                G_RC(g, R, grk_let_scatter(R.c, R.let_recv, loff[l][W]));
            }
        } else {
            // packed walk nodes of every rank's range: internal nodes [g0, g1) (ids < N-1) and leaves N-1+[g0, g1)
            g->last_let = -1;
            std::vector<int64_t> oi(W + 1), ol(W + 1);
            for (int r = 0; r <= W; r++) { oi[r] = std::min<int64_t>(g->g0[r], N - 1); ol[r] = N - 1 + g->g0[r]; }
            { auto b = ptrs(g, [](GroupRank& R) { return R.c->packed; }); if ((rc = g_allgatherv(g, b.data(), oi.data(), 32))) return rc; }
            { auto b = ptrs(g, [](GroupRank& R) { return R.c->packed; }); if ((rc = g_allgatherv(g, b.data(), ol.data(), 32))) return rc; }
        }
        aux_mark(g, "gravity_aux_node_exchange");
        FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_RC(g, R, sph_launch_top_tree(R.c, W, R.bnd, R.split_d, dt)); }
        aux_mark(g, "gravity_aux_top_tree");
        return SPH_OK;
    };
    // 3. local stable sort of the received set
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        R.n_own = roff[l][W];
        G_RC(g, R, grk_mig_keys(c, R.mig_recv, (int)R.n_own, c->keys[1]));
        G_RC(g, R, sph_launch_radix_sort(c, (int)R.n_own, c->stream));
        if (l == 0) pass_mark(g, "sort");
    }
    // 3b. the own particles are final: the gravity lane (source / key all-gathers, LBVH build of the own range, selection of the
    // walk nodes the other ranks can reach) starts here, on the auxiliary stream and communicator, behind the halo exchange and
    // the neighbor pass
    const bool overlap = impl != SPH_GRAVITY_NONE && aux_available(g);
    if (overlap) {
        if ((rc = aux_begin(g))) return rc;
        const int rcg = gravity_part_a();
        if ((rc = aux_pause(g))) return rc;     // always restores the main stream / communicator
        if (rcg) return rcg;
    }
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        // 4. halo lists
        G_RC(g, R, grk_halo_lists(c, c->keys[1], (int)R.n_own, R.split_d, W, R.rank, R.bmask, R.hmask, R.hcnt, R.htot, R.hlist));
        G_CUDA(g, cudaMemcpyAsync(R.cnt_d + (size_t)R.rank * W, R.htot, W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    }
    pass_mark(g, "halo_lists");
    { auto b = ptrs(g, [](GroupRank& R) { return R.cnt_d; }); if ((rc = g_allgather(g, b.data(), W * sizeof(uint32_t)))) return rc; }
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        G_CUDA(g, cudaMemcpyAsync(R.cnt_h, R.cnt_d, (size_t)W * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, R.c->stream));
    }
    if ((rc = sync_all(g))) return rc;                                   // ---- host sync 2: halo counts
    const uint32_t* H = g->r[0].cnt_h;
    std::vector<std::vector<int64_t>> hoff(g->nlocal, std::vector<int64_t>(W + 1, 0));      // halo-buffer offsets per source
    std::vector<std::vector<int64_t>> eoff(g->nlocal, std::vector<int64_t>(W + 1, 0));      // extended-slot offsets per source
    g->last_halo = 0;
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        const int me = R.rank;
        int64_t low = 0, nsend = 0;
        for (int p = 0; p < W; p++) { hoff[l][p + 1] = hoff[l][p] + H[p * W + me]; if (p < me) low += H[p * W + me]; nsend += H[me * W + p]; }
        const int64_t high = hoff[l][W] - low;
        for (int p = 0; p < W; p++) eoff[l][p] = p <= me ? hoff[l][p] : hoff[l][p] + R.n_own;
        R.own0 = low; R.n_ext = low + R.n_own + high; R.n_hsend = nsend;
        g->last_halo += low + high;
        if (low > g->cap_halo || high > g->cap_halo || nsend > 2 * g->cap_halo)
            G_FAIL(g, SPH_ERR_CAPACITY, "halo of rank " + std::to_string(me) + " (" + std::to_string(low) + " + " + std::to_string(high) + " in, " +
                                            std::to_string(nsend) + " out) exceeds the halo capacity " + std::to_string(g->cap_halo));
    }
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_RC(g, R, grk_halo_pack(c, R.mig_recv, c->idx[1], c->keys[1], R.hlist, (int)R.n_hsend, R.halo_send));
    }
    {
        auto s = ptrs(g, [](GroupRank& R) { return R.halo_send; });
        auto d = ptrs(g, [](GroupRank& R) { return R.halo_recv; });
        if ((rc = g_alltoallv(g, s.data(), d.data(), hoff, H, 32))) return rc;
    }
    const sph_GridParams& grid = *g->r[0].grid_h;
    const size_t ncell = (size_t)1 << (3 * std::min(std::max(grid.bits, 0), 8));
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_RC(g, R, grk_assemble_ext(c, R.mig_recv, c->idx[1], c->keys[1], R.halo_recv, (int)R.own0, (int)R.n_own, (int)R.n_ext, c->keys[0], ncell));
        c->n = R.n_ext; c->t0 = R.own0; c->t1 = R.own0 + R.n_own; c->row_base = R.own0;
        c->tree_off = R.own0 - g->g0[R.rank];   // resident slot of global slot s (the walk's targets)
        c->skeys = c->keys[0];
        c->cur = 0;
        c->resident = true;
        c->h_bound = grid.hmax;   // the global maximum of this step (all-reduced bounds): exact, covers the halo too
    }
    pass_mark(g, "halo_exchange_cells");

    // 5. neighbor rows + density + EOS for the own targets; (m/rho)P of the halo follows
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_RC(g, R, sph_launch_neighbors_density(c));
        // hlist indexes the own particles: shift the source by own0
        G_RC(g, R, grk_gather_f32(c, c->cvol + R.own0, R.hlist, (int)R.n_hsend, R.cv_send));
    }
    {
        auto s = ptrs(g, [](GroupRank& R) { return R.cv_send; });
        auto d = ptrs(g, [](GroupRank& R) { return R.c->cvol; });
        if ((rc = g_alltoallv(g, s.data(), d.data(), eoff, H, 4))) return rc;
    }
    pass_mark(g, "neighbors_density_eos");

    // 6. pressure gradient (independent of gravity: it runs first, more cover for the gravity lane)
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_RC(g, R, sph_launch_pressure(R.c)); }
    pass_mark(g, "pressure_grad");

    // 6b. gravity of the own targets
    if (impl == SPH_GRAVITY_NONE) {
        FOR_RANKS(g, R) {
            G_CUDA(g, cudaSetDevice(R.device));
            sphb200_ctx* c = R.c;
            G_CUDA(g, cudaMemsetAsync(c->grav, 0, (size_t)c->n * sizeof(float4), c->stream));
            G_CUDA(g, cudaMemsetAsync(c->npart, 0, (size_t)c->n * sizeof(int32_t), c->stream));
            G_CUDA(g, cudaMemsetAsync(c->napprox, 0, (size_t)c->n * sizeof(int32_t), c->stream));
        }
    } else {
        if (overlap) {
            if ((rc = aux_resume(g))) return rc;
            const int rcg = gravity_part_b();
            if ((rc = aux_end(g))) return rc;
            if (rcg) return rcg;
            if ((rc = aux_join(g))) return rc;
        } else {
            if ((rc = gravity_part_a())) return rc;
            if ((rc = gravity_part_b())) return rc;
        }
        pass_mark(g, "gravity_exchange_wait");   // what the source / tree exchange takes beyond the neighbor pass it runs behind
        FOR_RANKS(g, R) {
            G_CUDA(g, cudaSetDevice(R.device));
            if (R.n_own <= 0) continue;
            if (impl == SPH_GRAVITY_PARTICLE) G_RC(g, R, sph_launch_gravity_allpairs(R.c));
            else G_RC(g, R, sph_launch_tree_walk(R.c));
            G_RC(g, R, sph_launch_gravity_near(R.c, impl));
        }
    }
    pass_mark(g, gname);

    // 7. integration
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_RC(g, R, sph_launch_integrate(R.c, dt)); }
    pass_mark(g, "integrate");
    g->stepped = true;
    g->steps++;
    return SPH_OK;
}

static int group_errflags(sphb200_group* g) {
    int worst = SPH_OK;
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_CUDA(g, cudaMemcpyAsync(c->err_h, c->err_d, ERR_SLOTS * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        G_CUDA(g, cudaStreamSynchronize(c->stream));
        if (c->err_h[ERR_TOP_TREE]) { g->err = "top-tree list overflow (more than SPH_TOP_CAP nodes straddle a rank boundary)"; worst = SPH_ERR_TREE_STACK; }
        else if (c->err_h[ERR_TREE_STACK]) { g->err = "LBVH traversal stack overflow"; worst = SPH_ERR_TREE_STACK; }
        else if (c->err_h[ERR_NEIGHBOR_OVERFLOW] && worst == SPH_OK) {
            g->err = "neighbor list overflow: a particle has " + std::to_string(c->err_h[ERR_NEIGHBOR_OVERFLOW]) + " neighbors > max_neighbors=" +
                     std::to_string(g->p.max_neighbors) + " (lists truncated)";
            worst = SPH_ERR_NEIGHBOR_OVERFLOW;
        }
    }
    return worst;
}

int sphb200_group_sync(sph_group g) {
    if (!g) return SPH_ERR_INVALID_ARG;
    return group_errflags(g);
}

// ---- download: results return to the rank that holds the particle's body-order slice (over NVLink), then one DMA per field ----
static int redistribute_results(sphb200_group* g) {
    if (g->res_valid) return SPH_OK;
    const int W = g->world;
    int rc;
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        G_RC(g, R, grk_body_dest(c, c->orig[0] + R.own0, (int)R.n_own, g->chunk, R.dest));
        G_CUDA(g, cudaMemsetAsync(R.tot256, 0, 256 * sizeof(uint32_t), c->stream));
        G_RC(g, R, sph_launch_digit_pass(c, nullptr, nullptr, R.dest, (int)R.n_own, 0, nullptr, R.perm, R.tot256, c->stream));
        G_CUDA(g, cudaMemcpyAsync(R.cnt_d + (size_t)R.rank * W, R.tot256, W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    }
    { auto b = ptrs(g, [](GroupRank& R) { return R.cnt_d; }); if ((rc = g_allgather(g, b.data(), W * sizeof(uint32_t)))) return rc; }
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        G_CUDA(g, cudaMemcpyAsync(R.cnt_h, R.cnt_d, (size_t)W * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, R.c->stream));
    }
    if ((rc = sync_all(g))) return rc;
    const uint32_t* M = g->r[0].cnt_h;
    std::vector<std::vector<int64_t>> roff(g->nlocal, std::vector<int64_t>(W + 1, 0));
    for (int l = 0; l < g->nlocal; l++) {
        GroupRank& R = g->r[l];
        for (int p = 0; p < W; p++) roff[l][p + 1] = roff[l][p] + M[p * W + R.rank];
        R.n_res = roff[l][W];
        if (R.n_res != R.nbody) G_FAIL(g, SPH_ERR_STATE, "internal: result redistribution lost particles");
    }
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_RC(g, R, grk_result_pack(R.c, R.perm, (int)R.n_own, (int)R.own0, (float4*)R.mig_send)); }
    {
        auto s = ptrs(g, [](GroupRank& R) { return R.mig_send; });
        auto d = ptrs(g, [](GroupRank& R) { return R.mig_recv; });
        if ((rc = g_alltoallv(g, s.data(), d.data(), roff, M, 96))) return rc;
    }
    g->res_valid = true;
    return SPH_OK;
}

int sphb200_group_download(sph_group g, int field, void* dst, int stride) {
    if (!g) return SPH_ERR_INVALID_ARG;
    if (!g->resident) G_FAIL(g, SPH_ERR_STATE, "no particles uploaded");
    if (field < 0 || field >= SPH_FIELD_COUNT_) G_FAIL(g, SPH_ERR_INVALID_ARG, "unknown field");
    if (g->n_total == 0) return SPH_OK;
    if (!dst) G_FAIL(g, SPH_ERR_INVALID_ARG, "null dst");
    int rc = redistribute_results(g);
    if (rc) return rc;
    static const int words[SPH_FIELD_COUNT_ + 1] = {3, 3, 1, 2, 1, 1, 3, 6, 1, 7};
    const bool sm_record = field == SPH_FIELD_SMOOTHING && stride == (int)sizeof(sph_ParticleSmoothing);
    const int kf = sm_record ? (int)SPH_FIELD_COUNT_ : field;
    const int eb = words[kf] * 4;
    const bool direct = stride == eb && (field != SPH_FIELD_SMOOTHING || sm_record);
    if (stride < (field == SPH_FIELD_SMOOTHING ? 4 : field == SPH_FIELD_GRAVITY ? 16 : eb)) G_FAIL(g, SPH_ERR_INVALID_ARG, "stride too small");
    int64_t pb0 = 0;
    sphb200_group_body_range(g, g->n_total, &pb0, nullptr);
    FOR_RANKS(g, R) {
        G_CUDA(g, cudaSetDevice(R.device));
        sphb200_ctx* c = R.c;
        if (R.nbody == 0) continue;
        G_RC(g, R, grk_result_field(c, (const float4*)R.mig_recv, (int)R.n_res, kf, R.body0, (float*)c->stage_d));
        char* d = (char*)dst + (size_t)(R.body0 - pb0) * stride;
        G_CUDA(g, cudaMemcpyAsync(direct ? (void*)d : c->stage_h, c->stage_d, (size_t)R.nbody * eb, cudaMemcpyDeviceToHost, c->stream));
    }
    rc = group_errflags(g);   // also synchronises
    if (direct) return rc;
    FOR_RANKS(g, R) {
        const char* src = (const char*)R.c->stage_h;
        char* d = (char*)dst + (size_t)(R.body0 - pb0) * stride;
        const int64_t n = R.nbody;
        if (field == SPH_FIELD_SMOOTHING) {
            const bool full = stride >= (int)sizeof(sph_ParticleSmoothing);
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
        } else if (field == SPH_FIELD_GRAVITY) {
            const int w = stride >= 24 ? 24 : 16;
            for (int64_t i = 0; i < n; i++) memcpy(d + (size_t)i * stride, src + 24 * i, w);
        } else {
            for (int64_t i = 0; i < n; i++) memcpy(d + (size_t)i * stride, src + (size_t)i * eb, eb);
        }
    }
    return rc;
}

// out[0..11] as sphb200_diagnostics, reduced over the whole group
int sphb200_group_diagnostics(sph_group g, double* out12) {
    if (!g || !out12) return SPH_ERR_INVALID_ARG;
    if (!g->resident) G_FAIL(g, SPH_ERR_STATE, "no particles uploaded");
    for (int k = 0; k < 12; k++) out12[k] = 0;
    if (g->n_total == 0) return SPH_OK;
    int rc;
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_RC(g, R, sph_launch_diagnostics_range(R.c, (int)R.own0, (int)R.n_own)); }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->diag_d; }); if ((rc = g_allreduce(g, b.data(), 11, GR_F64, GR_SUM))) return rc; }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->diag_d + 12; }); if ((rc = g_allreduce(g, b.data(), 1, GR_U64, GR_MAX))) return rc; }
    GroupRank& R = g->r[0];
    G_CUDA(g, cudaSetDevice(R.device));
    G_CUDA(g, cudaMemcpyAsync(R.diag_h, R.c->diag_d, 16 * sizeof(double), cudaMemcpyDeviceToHost, R.c->stream));
    if ((rc = sync_all(g))) return rc;
    for (int k = 0; k < 10; k++) out12[k] = R.diag_h[k];
    out12[10] = R.diag_h[10] / (double)g->n_total;
    unsigned long long mx; memcpy(&mx, &R.diag_h[12], 8);
    out12[11] = (double)(mx & 0xffffffffull);
    return SPH_OK;
}

// (min, max, mean) x (rho, P, |grad Phi|, K rho) over the whole group
int sphb200_group_field_stats(sph_group g, double* out12) {
    if (!g || !out12) return SPH_ERR_INVALID_ARG;
    if (!g->resident) G_FAIL(g, SPH_ERR_STATE, "no particles uploaded");
    for (int k = 0; k < 12; k++) out12[k] = 0;
    if (g->n_total == 0) return SPH_OK;
    if (!g->stepped) G_FAIL(g, SPH_ERR_STATE, "no fields yet: run a step first");
    int rc;
    FOR_RANKS(g, R) { G_CUDA(g, cudaSetDevice(R.device)); G_RC(g, R, sph_launch_field_stats_range(R.c, (int)R.own0, (int)R.n_own)); }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->diag_d + 16; }); if ((rc = g_allreduce(g, b.data(), 4, GR_F64, GR_SUM))) return rc; }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->diag_d + 20; }); if ((rc = g_allreduce(g, b.data(), 4, GR_U32, GR_MIN))) return rc; }
    { auto b = ptrs(g, [](GroupRank& R) { return R.c->diag_d + 22; }); if ((rc = g_allreduce(g, b.data(), 4, GR_U32, GR_MAX))) return rc; }
    GroupRank& R = g->r[0];
    G_CUDA(g, cudaSetDevice(R.device));
    G_CUDA(g, cudaMemcpyAsync(R.diag_h + 16, R.c->diag_d + 16, 8 * sizeof(double), cudaMemcpyDeviceToHost, R.c->stream));
    if ((rc = sync_all(g))) return rc;
    sph_finish_field_stats(R.diag_h + 16, (const uint32_t*)(R.diag_h + 20), g->n_total, out12);
    return SPH_OK;
}

// Snapshot of this process's body slice.  One process driving the whole group writes one complete file (loadable by a single
// handle too); one process per GPU writes `path`.rNNN per rank.
static std::string snap_path(sph_group g, const char* path) {
    if (g->nlocal == g->world) return path;
    char suf[16]; snprintf(suf, sizeof(suf), ".r%03d", g->rank0);
    return std::string(path) + suf;
}

int sphb200_group_snapshot_save(sph_group g, const char* path) {
    if (!g || !path) return SPH_ERR_INVALID_ARG;
    if (!g->resident) G_FAIL(g, SPH_ERR_STATE, "no particles uploaded");
    int64_t b0 = 0, cnt = 0;
    sphb200_group_body_range(g, g->n_total, &b0, &cnt);
    const size_t n = (size_t)cnt;
    std::vector<float> pos(3 * n), vel(3 * n), mass(n), h(n);
    std::vector<int32_t> no(n);
    std::vector<sph_ParticleSmoothing> sm(n);
    int rc;
    // collective: every process takes part in the redistribution even if its slice is empty
    if ((rc = sphb200_group_download(g, SPH_FIELD_TRANSLATION, pos.data() ? (void*)pos.data() : (void*)&b0, 12)) && rc != SPH_ERR_NEIGHBOR_OVERFLOW) return rc;
    if ((rc = sphb200_group_download(g, SPH_FIELD_VELOCITY, vel.data() ? (void*)vel.data() : (void*)&b0, 12)) && rc != SPH_ERR_NEIGHBOR_OVERFLOW) return rc;
    if ((rc = sphb200_group_download(g, SPH_FIELD_MASS, mass.data() ? (void*)mass.data() : (void*)&b0, 4)) && rc != SPH_ERR_NEIGHBOR_OVERFLOW) return rc;
    if ((rc = sphb200_group_download(g, SPH_FIELD_SMOOTHING, sm.data() ? (void*)sm.data() : (void*)&b0, (int)sizeof(sph_ParticleSmoothing))) && rc != SPH_ERR_NEIGHBOR_OVERFLOW) return rc;
    for (size_t i = 0; i < n; i++) { h[i] = sm[i].influenceArea; no[i] = sm[i].neighbors; }
    return sph_snapshot_write(snap_path(g, path).c_str(), g->n_total, b0, cnt, g->steps, pos.data(), vel.data(), mass.data(), h.data(), no.data(), g->err);
}

// Loads this process's body slice from a complete snapshot at `path` (written by a single handle or a one-process group) or,
// failing that, from its own part `path`.rNNN.  n_total comes from the file.
int sphb200_group_snapshot_load(sph_group g, const char* path) {
    if (!g || !path) return SPH_ERR_INVALID_ARG;
    int64_t n_total = 0;
    {   // header only: the size of the whole set
        std::vector<float> a, b, c; std::vector<sph_ParticleSmoothing> d;
        int rc = sph_snapshot_read(path, 0, 0, &n_total, a, b, c, d, g->err);
        if (rc) rc = sph_snapshot_read(snap_path(g, path).c_str(), 0, -1, &n_total, a, b, c, d, g->err);
        if (rc) return rc;
    }
    if (n_total > g->cap_total) G_FAIL(g, SPH_ERR_CAPACITY, "the snapshot exceeds the group's capacity");
    int64_t b0 = 0, cnt = 0;
    sphb200_group_body_range(g, n_total, &b0, &cnt);
    std::vector<float> pos, vel, mass; std::vector<sph_ParticleSmoothing> sm;
    int rc = sph_snapshot_read(path, b0, cnt, &n_total, pos, vel, mass, sm, g->err);
    if (rc) rc = sph_snapshot_read(snap_path(g, path).c_str(), b0, cnt, &n_total, pos, vel, mass, sm, g->err);
    if (rc) return rc;
    int64_t dummy = 0;
    return sphb200_group_upload(g, n_total, cnt ? (void*)pos.data() : (void*)&dummy, 12, cnt ? (void*)vel.data() : (void*)&dummy, 12,
                                cnt ? (void*)mass.data() : (void*)&dummy, 4, cnt ? (void*)sm.data() : (void*)&dummy, (int)sizeof(sph_ParticleSmoothing));
}

int sphb200_group_info(sph_group g, sph_GroupInfo* out) {
    if (!g || !out) return SPH_ERR_INVALID_ARG;
    memset(out, 0, sizeof(*out));
    out->world = g->world; out->nlocal = g->nlocal; out->rank0 = g->rank0; out->transport = g->local ? 1 : (g->world > 1 ? 0 : 2);
    out->n_total = g->n_total; out->steps = g->steps; out->migrated_last_step = g->last_migrated; out->halo_last_step = g->last_halo; out->tree_nodes_last_step = g->last_let;
    out->cap_own = g->cap_own; out->cap_halo = g->cap_halo;
    for (int l = 0; l < g->nlocal && l < 32; l++) { out->n_own[l] = g->r[l].n_own; out->n_halo[l] = g->r[l].n_ext - g->r[l].n_own; }
    int64_t launches = 0;
    for (auto& R : g->r) launches += R.c->launches;
    out->launches = launches;
    return SPH_OK;
}

int sphb200_group_enable_timing(sph_group g, int enable) {
    if (!g) return SPH_ERR_INVALID_ARG;
    g->timing = enable != 0; g->npass = 0;
    return SPH_OK;
}

int sphb200_group_get_timings(sph_group g, const char** names, float* ms, int cap) {
    if (!g) return SPH_ERR_INVALID_ARG;
    if (!g->timing || g->npass == 0) return 0;
    cudaSetDevice(g->r[0].device);
    cudaEventSynchronize(g->tev[g->npass]);
    int k = 0;
    for (; k < g->npass && k < cap; k++) {
        if (names) names[k] = g->pass_name[k];
        float t = 0;
        cudaEventElapsedTime(&t, g->tev[k], g->tev[k + 1]);
        if (ms) ms[k] = t;
    }
    // auxiliary lane: time of every stage since the previous one (the first: since the fork point on the main lane)
    for (int a = 0; a < g->naux && k < cap; a++, k++) {
        if (names) names[k] = g->aux_name[a];
        float t = 0;
        cudaEventSynchronize(g->aev[a + 1]);
        cudaEventElapsedTime(&t, a == 0 ? g->tev[g->aux_fork_pass] : g->aev[a], g->aev[a + 1]);
        if (ms) ms[k] = t;
    }
    return k;
}

int sphb200_group_stream(sph_group g, int local_rank, void** s) {
    if (!g || !s || local_rank < 0 || local_rank >= g->nlocal) return SPH_ERR_INVALID_ARG;
    *s = (void*)g->r[local_rank].c->stream;
    return SPH_OK;
}

// the context of local rank l (stage-level inspection in tests: device_ptr, download_sort, ...); owned by the group
int sphb200_group_rank_handle(sph_group g, int local_rank, sph_handle* out) {
    if (!g || !out || local_rank < 0 || local_rank >= g->nlocal) return SPH_ERR_INVALID_ARG;
    *out = g->r[local_rank].c;
    return SPH_OK;
}

}  // extern "C"
