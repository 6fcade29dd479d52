/* sphb200.h -- C ABI of the B200-native SPH hot path (drop-in for PlanetModel-SPH's per-timestep systems).
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes, returns an int status (0 = ok, <0 = error)
 * and never exposes torch or CUDA types.  Each one cites the reference interface it replaces
 * (A/ = Assets/Scripts/, UP/ = UpstreamPackages/com.unity.physics@0.6.0-preview.3/Unity.Physics/).
 * The C# P/Invoke binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Threading: one host thread per handle; calls are asynchronous on the handle's CUDA stream except
 * upload/download/sync/diagnostics which block.  Distinct handles are independent.
 * Ownership: the library owns all device memory; host pointers are borrowed for the duration of a call.
 * There is NO CPU fallback: every call fails with SPH_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SPHB200_H
#define SPHB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SPH_API __declspec(dllexport)
#else
#define SPH_API __attribute__((visibility("default")))
#endif

/* ---- status codes (reference: C# exceptions, KernelSystem.cs:102-105, GravityFieldSystem.cs:84-87,400-403) */
enum {
    SPH_OK = 0,
    SPH_ERR_INVALID_ARG = -1,
    SPH_ERR_CAPACITY = -2,          /* n > capacity (the reference itself is capped at 2^24-2 bodies, Scheduler.cs:26-31,41; this library is not) */
    SPH_ERR_NEIGHBOR_OVERFLOW = -3, /* some particle has more than max_neighbors neighbors; lists truncated.  Raised asynchronously: reported by
                                       the next download / sync while the lists in memory are the truncated ones; the next neighbor pass whose
                                       rows fit clears it (density and the h controller stay complete under overflow, so a run recovers by itself) */
    SPH_ERR_CUDA = -4,
    SPH_ERR_STATE = -5,             /* stage called out of order (e.g. pressure before build_neighbors) */
    SPH_ERR_TREE_STACK = -6,        /* traversal stack / top-tree list overflow in the LBVH gravity */
    SPH_ERR_NCCL = -7               /* multi-GPU group: libnccl missing or a collective failed */
};

/* ---- gravity implementation (GravityFieldSystem.cs:19-25 `GravityImpl`) */
enum {
    SPH_GRAVITY_TREE = 0,     /* GRAVITY_TREE_CPU: Barnes-Hut monopole, Bmax MAC (default in the reference) */
    SPH_GRAVITY_PARTICLE = 1, /* GRAVITY_PARTICLE_CPU: O(N^2) direct sum */
    SPH_GRAVITY_NONE = 2
};

/* ---- compile-time constants of the reference, as one POD (SURVEY.md section 5 "Config / flags") */
typedef struct sph_Params {
    float K;                /* EOS constant P = K rho^2           PressureFieldSystem.cs:31 (1000)  */
    float G;                /* k_GravConstant                     GravityFieldSystem.cs:26  (1)     */
    float theta;            /* k_Theta                            GravityFieldSystem.cs:228 (0.7)   */
    float target_neighbors; /* TARGET_NEIGHBORS                   ParticleSmoothingSystem.cs:18 (50) */
    int32_t max_neighbors;  /* neighbor-list capacity per particle (multiple of 32; default 256)    */
    int32_t leaf_max;       /* bodies per tree leaf               BoundingVolumeHierarchy.cs:40-83 (4) */
    int32_t aabb_mode;      /* 0 = MAC boxes are collider AABBs (quirk Q2, default), 1 = point bounds */
    int32_t max_grid_bits;  /* cells per axis <= 2^bits, bits <= 8; 0 = choose from capacity          */
    int32_t flags;          /* SPH_FLAG_* ; 0 = reference-compatible                                  */
    int32_t reserved[3];
} sph_Params;

enum {
    SPH_FLAG_FIX_KERNEL_DERIV_SIGN = 1, /* use -3q in KernelDeriv's inner branch (undo quirk Q1, SplineKernel.cs:135) */
    SPH_FLAG_KICK_DRIFT = 2,            /* v += a dt first, then x += v_new dt (symplectic leapfrog; roadmap README.md:90-93) */
    SPH_FLAG_PM07_SOFTENING = 4         /* gravity softened with the spline kernel, symmetric in h_i and h_j (Price & Monaghan 2007;
                                           roadmap README.md:75-77) instead of the one-sided a = h_i law of GravityFieldSystem.cs:340-347 */
};

/* ---- byte-exact mirrors of the reference components */
typedef struct { float x, y, z; } sph_Translation;                        /* Unity.Transforms.Translation, 12 B */
typedef struct { float linear[3]; float angular[3]; } sph_PhysicsVelocity; /* UP/ECS/Base/Components/PhysicsComponents.cs:79-90, 24 B */
typedef struct { float value; } sph_ParticleMass;                          /* A/Components/DensityField.cs:4-7   */
typedef struct { float value; } sph_ParticleDensity;                       /* A/Components/DensityField.cs:9-13  */
typedef struct { float value; } sph_ParticlePressure;                      /* A/Components/PressureField.cs:4-7  */
typedef struct { float value[3]; } sph_ParticlePressureGrad;               /* A/Components/PressureField.cs:9-12 */
typedef struct {
    float influenceArea;              /* h   */
    float supportDomain;              /* 2h  */
    float sphereColliderPosRadius[4]; /* debug: xyz = centre, w = radius (= 2h) */
    int32_t neighbors;                /* debug: own-support neighbor count */
} sph_ParticleSmoothing;                                                   /* A/Components/ParticleSmoothing.cs:28-31, 28 B */
typedef struct {
    float value[4];       /* xyz = grad(Phi), w = Phi */
    int32_t numParticles; /* direct body interactions in the tree walk */
    int32_t numApprox;    /* accepted node approximations */
} sph_GravityField;                                                        /* A/Components/GravityField.cs:10-15, 24 B */
typedef struct {
    int32_t otherIndex, otherVersion; /* Entity Other */
    float kernelThis[4];              /* xyz = grad_i W(r,h_i), w = W(r,h_i) */
    float kernelSymmetric[4];
} sph_ParticleInteraction;                                                 /* A/Components/Kernel.cs:6-10, 40 B */

/* Grid/key parameters of the last neighbor build (specification shared with oracle/sph_oracle.cpp GridParams) */
typedef struct sph_GridParams {
    float min[3];       /* global AABB min of the positions */
    float cell;         /* cell edge = max(2.002*href, 2.002*hmax/4, ext*1.0001/2^bits_max) */
    float fine_scale;   /* 1024 / (cell * 2^bits): positions -> 10-bit fine Morton coordinates */
    int32_t bits;       /* cells per axis = 2^bits */
    float hmax;
    float ext;          /* largest AABB extent */
    float href;         /* typical h: float whose bit pattern is the mean bit pattern of all h (deterministic) */
    int32_t stencil;    /* S: interacting pairs lie within S cells of each other on every axis (S*cell >= 2.002*hmax) */
} sph_GridParams;

/* ---- per-particle fields for download / device_ptr */
enum {
    SPH_FIELD_TRANSLATION = 0,   /* sph_Translation           (natural stride 12) */
    SPH_FIELD_VELOCITY = 1,      /* sph_PhysicsVelocity.linear (12 B written; natural stride 24) */
    SPH_FIELD_MASS = 2,          /* 4  */
    SPH_FIELD_SMOOTHING = 3,     /* sph_ParticleSmoothing if stride >= 28, else h only (4 B) */
    SPH_FIELD_DENSITY = 4,       /* 4  */
    SPH_FIELD_PRESSURE = 5,      /* 4  */
    SPH_FIELD_PRESSURE_GRAD = 6, /* 12 */
    SPH_FIELD_GRAVITY = 7,       /* sph_GravityField if stride >= 24, else float4 (16 B) */
    SPH_FIELD_NEIGHBOR_COUNT = 8,/* int32: symmetric neighbor count (interaction buffer length) */
    SPH_FIELD_COUNT_
};

typedef struct sphb200_ctx* sph_handle;

/* Fill *p with the reference's constants. */
SPH_API int sphb200_default_params(sph_Params* p);

/* Create a simulation context on CUDA device `device` holding up to `capacity` particles.
 * Replaces: World/system creation (OnCreate of the six systems, e.g. GravityFieldSystem.cs:39-46). */
SPH_API int sphb200_create(const sph_Params* params, int64_t capacity, int device, sph_handle* out);
SPH_API int sphb200_destroy(sph_handle h);
/* Last error text of this handle (or of the failed create when h == NULL). */
SPH_API const char* sphb200_last_error(sph_handle h);
/* Run on an externally owned cudaStream_t (e.g. torch's current stream); NULL = the handle's own stream. */
SPH_API int sphb200_set_stream(sph_handle h, void* cuda_stream);
/* The cudaStream_t the handle currently runs on (for CUDA-event timing by the caller). */
SPH_API int sphb200_get_stream(sph_handle h, void** cuda_stream);
/* Blocks until the handle's stream is idle; returns sticky asynchronous errors of the steps issued since the last
 * neighbor build (SPH_ERR_NEIGHBOR_OVERFLOW, SPH_ERR_TREE_STACK). */
SPH_API int sphb200_sync(sph_handle h);

/* Upload the per-particle components from host AoS arrays (strides in bytes; body index = array index).
 * Replaces: BuildPhysicsWorld's entity->rigid-body gather (UP/ECS/Base/Systems/BuildPhysicsWorld.cs:389-469).
 *   pos        -> sph_Translation          (Translation.Value)
 *   vel        -> sph_PhysicsVelocity      (.linear read)
 *   mass       -> sph_ParticleMass
 *   smoothing  -> sph_ParticleSmoothing    (.influenceArea = h read; .neighbors read as last step's own-support count
 *                                           when stride >= 28, else 0)
 * Resets the resident order to body-index order. */
SPH_API int sphb200_upload(sph_handle h, int64_t n, const void* pos, int pos_stride, const void* vel, int vel_stride,
                           const void* mass, int mass_stride, const void* smoothing, int smoothing_stride);

/* ---- stages, one per reference system ------------------------------------------ */
/* ParticleSmoothingSystem.OnUpdate (ParticleSmoothingSystem.cs:19-87): h <- h*0.5*(1+(target/n)^(1/3)). */
SPH_API int sphb200_smoothing_update(sph_handle h);
/* KernelSystem.OnUpdate + BuildPhysicsWorld broadphase (KernelSystem.cs:93-231, Broadphase.cs:275-351):
 * keys, radix sort, cell table, neighbor lists (exact Interacts + keep rule), and -- fused -- the density sum. */
SPH_API int sphb200_build_neighbors(sph_handle h);
/* GravityFieldSystem.OnUpdate (GravityFieldSystem.cs:62-74): impl = SPH_GRAVITY_*.  dt is needed because the
 * reference's MAC boxes are swept by v*dt (quirk Q2). PARTICLE mode requires build_neighbors in this step. */
SPH_API int sphb200_gravity(sph_handle h, int impl, float dt);
/* Optional hint, called between smoothing_update and build_neighbors: announces the gravity path of this step so that
 * the LBVH build (which needs only the sorted particles) runs on an auxiliary stream concurrently with the neighbor
 * pass.  The reference builds its tree in BuildPhysicsWorld, i.e. also before KernelSystem (BuildPhysicsWorld.cs:286-289). */
SPH_API int sphb200_prepare_gravity(sph_handle h, int impl, float dt);
/* DensityFieldSystem.OnUpdate (DensityFieldSystem.cs:38-56): publishes rho (summed inside build_neighbors). */
SPH_API int sphb200_density(sph_handle h);
/* PressureFieldSystem.OnUpdate (PressureFieldSystem.cs:30-70): P = K rho^2 and grad P. */
SPH_API int sphb200_pressure(sph_handle h);
/* Integrator.IntegratePosition + VelocitySystem.OnUpdate (Integrator.cs:98-101, VelocitySystem.cs:24-36). */
SPH_API int sphb200_integrate(sph_handle h, float dt);
/* One FixedStepSimulationSystemGroup tick: all of the above in the reference order. */
SPH_API int sphb200_step(sph_handle h, float dt, int gravity_impl);

/* Restrict the *target* particles of density/pressure/gravity/integrate to sorted slots [t0,t1) (multi-GPU:
 * Morton-range ownership).  t1 < 0 = all.  Neighbor/tree structures are always built over all resident particles. */
SPH_API int sphb200_set_target_range(sph_handle h, int64_t t0, int64_t t1);

/* ---- results --------------------------------------------------------------------------------------------- */
/* Download one field into a host AoS array in body-index order.
 * Replaces: ExportPhysicsWorld (UP/ECS/Base/Systems/ExportPhysicsWorld.cs:130-161) and the component writes. */
SPH_API int sphb200_download(sph_handle h, int field, void* dst, int stride);
/* Neighbor lists as CSR in body-index space, each row ascending (canonical order).  offsets has n+1 entries.
 * If total > cap only *total is set (call again).  Replaces reading DynamicBuffer<ParticleInteraction>.Other. */
SPH_API int sphb200_download_neighbors(sph_handle h, int64_t* offsets, int32_t* nbr, int64_t cap, int64_t* total);
/* Interaction records (KernelSystem.cs:305-334) for the same CSR layout; optional debugging/parity surface.
 * SPH_ERR_STATE unless the lists of build_neighbors still describe the resident positions (integrate / smoothing_update
 * invalidate them); SPH_ERR_INVALID_ARG for offsets that do not ascend from 0 or an index outside [0, n). */
SPH_API int sphb200_download_interactions(sph_handle h, const int64_t* offsets, const int32_t* nbr,
                                          sph_ParticleInteraction* out);
/* Sorted slot -> body index, and the 30-bit Morton keys in sorted order (either pointer may be NULL). */
SPH_API int sphb200_download_sort(sph_handle h, uint32_t* order, uint32_t* keys, sph_GridParams* grid);
/* LBVH of the last tree-gravity call: 2n-1 nodes (internal 0..n-2, leaf slot s -> n-1+s). Any pointer may be NULL.
 * child = int32[2] per node, range = int32[2] (first,last), moment = float[4] (cm, M), lo/hi = float[3]. */
SPH_API int sphb200_download_tree(sph_handle h, int32_t* child, int32_t* range, float* moment, float* lo, float* hi);
/* Conserved-quantity diagnostics (README.md:50-52 roadmap): out[0]=sum m, [1..3]=sum m v, [4..6]=sum m x cross v,
 * [7]=E_kin, [8]=E_pot=0.5 sum m Phi, [9]=E_int=sum m K rho, [10]=mean symmetric neighbor count, [11]=max count */
SPH_API int sphb200_diagnostics(sph_handle h, double* out12);
/* SPH_DEBUG_BOUNDS build (make -C csrc debug -> libsphb200_dbg.so; the release library answers *allocations = -1): every device
 * allocation of the library carries guard zones, every data-dependent index in the kernels is asserted (a violation traps and the
 * next synchronising call fails).  *bad_bytes = guard bytes overwritten so far, over *allocations live allocations.  Stands in
 * for the reference's job-safety checks (KernelSystem.cs:247, 475, 546 [NativeDisableContainerSafetyRestriction] opt-outs). */
SPH_API int sphb200_debug_check_guards(int64_t* bad_bytes, int64_t* allocations);

/* Field statistics (README.md:50-52 roadmap: "Average/max/min: Temp, Pressure, Density, Grav Field"): for q = rho, P,
 * |grad Phi|, u = K rho (specific internal energy of the P = K rho^2 gas, the model's temperature proxy), in this order:
 * out[3q] = min, out[3q+1] = max, out[3q+2] = mean over all particles. */
SPH_API int sphb200_field_stats(sph_handle h, double* out12);
/* Snapshot dump / restore of the host-visible state (positions, velocities, masses, smoothing lengths, own-support counts)
 * in body order: checkpoint / resume, and the format fixtures travel in.  One binary file, 64-byte header.  Restoring and
 * stepping continues the run bit-identically (the resident order is rebuilt from the body order by the next sort). */
SPH_API int sphb200_snapshot_save(sph_handle h, const char* path);
SPH_API int sphb200_snapshot_load(sph_handle h, const char* path);

/* ---- introspection ------------------------------------------------------------------------------------------ */
SPH_API int sphb200_count(sph_handle h, int64_t* n, int64_t* capacity);
SPH_API int sphb200_get_params(sph_handle h, sph_Params* out);
/* Raw device pointer of an internal SoA array (sorted order) for zero-copy interop (NCCL through torch).
 * names: "posh" float4(x,y,z,h), "velm" float4(v,m), "posm" float4(x,y,z,m), "rho", "press", "cvol" float,
 * "gradp" float4, "grav" float4, "nown" int32, "orig" uint32 */
SPH_API int sphb200_device_ptr(sph_handle h, const char* name, void** ptr, int64_t* bytes);
/* Number of kernel launches issued by this library since create (bench.py's gpu_launches). */
SPH_API int sphb200_launch_count(sph_handle h, int64_t* launches);
/* Per-pass device times (ms) of the most recent step when timing is enabled.
 * names[i] -> static strings; returns number of passes written (<= cap). */
SPH_API int sphb200_enable_timing(sph_handle h, int enable);
SPH_API int sphb200_get_timings(sph_handle h, const char** names, float* ms, int cap);
/* FP32 FMA-pipe microbenchmark (roofline denominator for all-pairs gravity): returns TFLOP/s. */
SPH_API int sphb200_fp32_peak(sph_handle h, double* tflops);
SPH_API const char* sphb200_version(void);

/* ---- multi-GPU groups: Morton-range domain decomposition with NVLink halo exchange -------------------------------------
 * The reference has one process and shared-memory job threads only (UP/Collision/World/Broadphase.cs:163) and is capped at
 * 2^24-2 bodies (UP/Dynamics/Simulation/Scheduler.cs:24-41); a group is how this library scales N.
 * Every rank (one per GPU) keeps the particles of one contiguous Morton-key range plus a halo; per step the ranks exchange
 * migrating particles, halo particles, (m/rho)P of the halo, and all-gather the gravity sources (and the packed LBVH
 * nodes for tree gravity) over NCCL.  Results are bit-identical to a single handle for tree gravity.
 * Body-order slices: with chunk = ceil(n/world), rank r uploads and downloads bodies [r*chunk, min((r+1)*chunk, n)):
 * host<->device traffic per GPU is 1/world of the state; the particles travel between GPUs over NVLink. */
typedef struct sphb200_group* sph_group;

typedef struct sph_GroupInfo {
    int32_t world, nlocal, rank0;
    int32_t transport;            /* 0 = NCCL, 1 = in-process peer copies, 2 = none (world 1) */
    int64_t n_total, steps;
    int64_t migrated_last_step;   /* particles that changed rank in the last step (this process's ranks, incoming) */
    int64_t halo_last_step;       /* halo particles received in the last step (this process's ranks) */
    int64_t cap_own, cap_halo;    /* per-rank capacities: own slots, halo slots on either side */
    int64_t launches;             /* kernels launched by this process's ranks since create */
    int64_t tree_nodes_last_step; /* walk-node records received in the last tree-gravity step (locally essential tree; this process's
                                     ranks); -1 = the whole node array was all-gathered (records beyond the exchange capacity) */
    int64_t n_own[32], n_halo[32];/* per local rank */
} sph_GroupInfo;

/* One process drives `ndev` GPUs (what a C# host does: INTEGRATION.md).  devices == NULL: 0..ndev-1.  NCCL communicators come
 * from ncclCommInitAll; if a device is listed twice (several ranks on one GPU: tests) or SPHB200_GROUP_TRANSPORT=local, the
 * ranks exchange by in-process peer copies instead.  capacity = particles of the whole group.
 * Replaces: World creation (OnCreate of the six systems), as sphb200_create. */
SPH_API int sphb200_group_create(const sph_Params* params, int64_t capacity, int ndev, const int* devices, sph_group* out);
/* One process per GPU (torchrun, MPI): rank 0 calls sphb200_group_unique_id, the host distributes the 128 bytes, every
 * process calls sphb200_group_create_rank with its rank and device. */
SPH_API int sphb200_group_unique_id(void* id128);
SPH_API int sphb200_group_create_rank(const sph_Params* params, int64_t capacity, const void* id128, int world, int rank,
                                      int device, sph_group* out);
SPH_API int sphb200_group_destroy(sph_group g);
SPH_API const char* sphb200_group_last_error(sph_group g);
/* Bodies this process uploads / downloads: [*body0, *body0 + *count). */
SPH_API int sphb200_group_body_range(sph_group g, int64_t n_total, int64_t* body0, int64_t* count);
/* As sphb200_upload; the arrays hold this process's bodies only (element 0 = body *body0). n_total = bodies of the group. */
SPH_API int sphb200_group_upload(sph_group g, int64_t n_total, const void* pos, int pos_stride, const void* vel, int vel_stride,
                                 const void* mass, int mass_stride, const void* smoothing, int smoothing_stride);
/* One FixedStepSimulationSystemGroup tick on all ranks (as sphb200_step). */
SPH_API int sphb200_group_step(sph_group g, float dt, int gravity_impl);
/* As sphb200_download, for this process's bodies (element 0 = body *body0). */
SPH_API int sphb200_group_download(sph_group g, int field, void* dst, int stride);
SPH_API int sphb200_group_sync(sph_group g);
/* As sphb200_diagnostics, reduced over the group. */
SPH_API int sphb200_group_diagnostics(sph_group g, double* out12);
/* As sphb200_field_stats, reduced over the group. */
SPH_API int sphb200_group_field_stats(sph_group g, double* out12);
/* As sphb200_snapshot_save / _load for this process's body slice.  A process that drives the whole group writes one complete
 * file (a single handle can load it, and vice versa); one process per GPU writes / reads its part `path`.rNNN, and also
 * accepts a complete file at `path` on load (every process reads its slice of it). */
SPH_API int sphb200_group_snapshot_save(sph_group g, const char* path);
SPH_API int sphb200_group_snapshot_load(sph_group g, const char* path);
SPH_API int sphb200_group_info(sph_group g, sph_GroupInfo* out);
SPH_API int sphb200_group_enable_timing(sph_group g, int enable);
SPH_API int sphb200_group_get_timings(sph_group g, const char** names, float* ms, int cap);
/* cudaStream_t of local rank `local_rank` (for CUDA-event timing by the caller). */
SPH_API int sphb200_group_stream(sph_group g, int local_rank, void** cuda_stream);
/* Context of local rank `local_rank` (owned by the group) for inspection through the sphb200_* getters. */
SPH_API int sphb200_group_rank_handle(sph_group g, int local_rank, sph_handle* out);

#ifdef __cplusplus
}
#endif
#endif /* SPHB200_H */
