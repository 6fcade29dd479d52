// SphB200Native.cs -- P/Invoke surface of libsphb200 (include/sphb200.h), the binding a PlanetModel-SPH maintainer
// adds to Assets/Scripts/.  Shipped as source: the build image has no C#/Unity toolchain, so this file is exercised
// only through its contract with the C header: tests/test_abi_exports.py compiles include/sphb200.h and checks every
// size/offset the [StructLayout] attributes below assume, and tests/test_csharp_binding.py parses this file against the
// header -- every SPH_API function has exactly one [DllImport] here with the same parameter list, every mirrored struct
// the same fields in the same order, every status / flag / field constant the same value.
using System;
using System.Runtime.InteropServices;

public static unsafe class SphB200Native
{
    const string Lib = "sphb200";   // libsphb200.so / sphb200.dll next to the player binary

    public const int SPH_OK = 0, SPH_ERR_INVALID_ARG = -1, SPH_ERR_CAPACITY = -2, SPH_ERR_NEIGHBOR_OVERFLOW = -3,
                     SPH_ERR_CUDA = -4, SPH_ERR_STATE = -5, SPH_ERR_TREE_STACK = -6, SPH_ERR_NCCL = -7;

    // sph_Params.flags (all off by default = the reference's behaviour): undo quirk Q1; kick-drift integration (README.md:90-93);
    // Price & Monaghan 2007 spline-softened, h-symmetric gravity (README.md:75-77)
    public const int SPH_FLAG_FIX_KERNEL_DERIV_SIGN = 1, SPH_FLAG_KICK_DRIFT = 2, SPH_FLAG_PM07_SOFTENING = 4;

    // GravityFieldSystem.GravityImpl (GravityFieldSystem.cs:19-23)
    public const int SPH_GRAVITY_TREE = 0, SPH_GRAVITY_PARTICLE = 1, SPH_GRAVITY_NONE = 2;

    public enum Field { Translation = 0, Velocity = 1, Mass = 2, Smoothing = 3, Density = 4, Pressure = 5, PressureGrad = 6,
                        Gravity = 7, NeighborCount = 8 }

    [StructLayout(LayoutKind.Sequential)]
    public struct Params
    {
        public float K, G, theta, target_neighbors;
        public int max_neighbors, leaf_max, aabb_mode, max_grid_bits, flags;
        public fixed int reserved[3];
    }

    // sph_GridParams: grid / key parameters of the last neighbor build (sphb200_download_sort)
    [StructLayout(LayoutKind.Sequential)]
    public struct GridParams
    {
        public fixed float min[3];
        public float cell, fine_scale;
        public int bits;
        public float hmax, ext, href;
        public int stencil;
    }

    // sph_ParticleInteraction = A/Components/Kernel.cs:6-10 (Entity Other; float4 KernelThis; float4 KernelSymmetric), 40 B
    [StructLayout(LayoutKind.Sequential)]
    public struct ParticleInteraction
    {
        public int otherIndex, otherVersion;
        public fixed float kernelThis[4];
        public fixed float kernelSymmetric[4];
    }

    // sph_GroupInfo (sphb200_group_info)
    [StructLayout(LayoutKind.Sequential)]
    public struct GroupInfo
    {
        public int world, nlocal, rank0, transport;
        public long n_total, steps, migrated_last_step, halo_last_step, cap_own, cap_halo, launches, tree_nodes_last_step;
        public fixed long n_own[32];
        public fixed long n_halo[32];
    }

    // ---- one handle = one GPU.  Order and grouping follow include/sphb200.h.
    [DllImport(Lib)] public static extern int sphb200_default_params(Params* p);
    [DllImport(Lib)] public static extern int sphb200_create(Params* p, long capacity, int device, out IntPtr handle);
    [DllImport(Lib)] public static extern int sphb200_destroy(IntPtr h);
    [DllImport(Lib)] public static extern IntPtr sphb200_last_error(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_set_stream(IntPtr h, void* cudaStream);
    [DllImport(Lib)] public static extern int sphb200_get_stream(IntPtr h, out IntPtr cudaStream);
    [DllImport(Lib)] public static extern int sphb200_sync(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_upload(IntPtr h, long n, void* pos, int posStride, void* vel, int velStride,
                                                            void* mass, int massStride, void* smoothing, int smoothingStride);
    // the six systems (SphB200Systems.cs makes exactly one of these calls per OnUpdate)
    [DllImport(Lib)] public static extern int sphb200_smoothing_update(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_build_neighbors(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_gravity(IntPtr h, int impl, float dt);
    [DllImport(Lib)] public static extern int sphb200_prepare_gravity(IntPtr h, int impl, float dt);
    [DllImport(Lib)] public static extern int sphb200_density(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_pressure(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_integrate(IntPtr h, float dt);
    [DllImport(Lib)] public static extern int sphb200_step(IntPtr h, float dt, int gravityImpl);
    [DllImport(Lib)] public static extern int sphb200_set_target_range(IntPtr h, long t0, long t1);
    // results
    [DllImport(Lib)] public static extern int sphb200_download(IntPtr h, int field, void* dst, int stride);
    [DllImport(Lib)] public static extern int sphb200_download_neighbors(IntPtr h, long* offsets, int* nbr, long cap, long* total);
    [DllImport(Lib)] public static extern int sphb200_download_interactions(IntPtr h, long* offsets, int* nbr, ParticleInteraction* records);
    [DllImport(Lib)] public static extern int sphb200_download_sort(IntPtr h, uint* order, uint* keys, GridParams* grid);
    [DllImport(Lib)] public static extern int sphb200_download_tree(IntPtr h, int* child, int* range, float* moment, float* lo, float* hi);
    [DllImport(Lib)] public static extern int sphb200_diagnostics(IntPtr h, double* out12);
    [DllImport(Lib)] public static extern int sphb200_debug_check_guards(long* badBytes, long* allocations);
    [DllImport(Lib)] public static extern int sphb200_field_stats(IntPtr h, double* out12);            // min / max / mean rho, P, |grad Phi|, u (README.md:50-52)
    [DllImport(Lib)] public static extern int sphb200_snapshot_save(IntPtr h, [MarshalAs(UnmanagedType.LPStr)] string path);
    [DllImport(Lib)] public static extern int sphb200_snapshot_load(IntPtr h, [MarshalAs(UnmanagedType.LPStr)] string path);
    // introspection
    [DllImport(Lib)] public static extern int sphb200_count(IntPtr h, long* n, long* capacity);
    [DllImport(Lib)] public static extern int sphb200_get_params(IntPtr h, Params* p);
    [DllImport(Lib)] public static extern int sphb200_device_ptr(IntPtr h, [MarshalAs(UnmanagedType.LPStr)] string name, out IntPtr ptr, long* bytes);
    [DllImport(Lib)] public static extern int sphb200_launch_count(IntPtr h, long* launches);
    [DllImport(Lib)] public static extern int sphb200_enable_timing(IntPtr h, int enable);
    [DllImport(Lib)] public static extern int sphb200_get_timings(IntPtr h, IntPtr* names, float* ms, int cap);   // names[i]: static ANSI strings
    [DllImport(Lib)] public static extern int sphb200_fp32_peak(IntPtr h, double* tflops);
    [DllImport(Lib)] public static extern IntPtr sphb200_version();

    // ---- several GPUs: Morton-range domain decomposition (include/sphb200.h "group" section).  One process -- the Unity player --
    // drives every GPU of the node through ONE group handle; body indices, strides and fields mean what they mean above.
    [DllImport(Lib)] public static extern int sphb200_group_create(Params* p, long capacity, int ndev, int* devices, out IntPtr group);
    // one process per GPU (a headless batch host under mpirun / torchrun): rank 0 makes the id, the host distributes its 128 bytes
    [DllImport(Lib)] public static extern int sphb200_group_unique_id(void* id128);
    [DllImport(Lib)] public static extern int sphb200_group_create_rank(Params* p, long capacity, void* id128, int world, int rank, int device, out IntPtr group);
    [DllImport(Lib)] public static extern int sphb200_group_destroy(IntPtr g);
    [DllImport(Lib)] public static extern IntPtr sphb200_group_last_error(IntPtr g);
    [DllImport(Lib)] public static extern int sphb200_group_body_range(IntPtr g, long nTotal, long* body0, long* count);
    [DllImport(Lib)] public static extern int sphb200_group_upload(IntPtr g, long nTotal, void* pos, int posStride, void* vel, int velStride,
                                                                  void* mass, int massStride, void* smoothing, int smoothingStride);
    [DllImport(Lib)] public static extern int sphb200_group_step(IntPtr g, float dt, int gravityImpl);
    [DllImport(Lib)] public static extern int sphb200_group_download(IntPtr g, int field, void* dst, int stride);
    [DllImport(Lib)] public static extern int sphb200_group_sync(IntPtr g);
    [DllImport(Lib)] public static extern int sphb200_group_diagnostics(IntPtr g, double* out12);
    [DllImport(Lib)] public static extern int sphb200_group_field_stats(IntPtr g, double* out12);
    [DllImport(Lib)] public static extern int sphb200_group_snapshot_save(IntPtr g, [MarshalAs(UnmanagedType.LPStr)] string path);
    [DllImport(Lib)] public static extern int sphb200_group_snapshot_load(IntPtr g, [MarshalAs(UnmanagedType.LPStr)] string path);
    [DllImport(Lib)] public static extern int sphb200_group_info(IntPtr g, GroupInfo* info);
    [DllImport(Lib)] public static extern int sphb200_group_enable_timing(IntPtr g, int enable);
    [DllImport(Lib)] public static extern int sphb200_group_get_timings(IntPtr g, IntPtr* names, float* ms, int cap);
    [DllImport(Lib)] public static extern int sphb200_group_stream(IntPtr g, int localRank, out IntPtr cudaStream);
    [DllImport(Lib)] public static extern int sphb200_group_rank_handle(IntPtr g, int localRank, out IntPtr handle);

    public static void CheckGroup(IntPtr g, int rc)
    {
        if (rc == SPH_OK) return;
        string msg = Marshal.PtrToStringAnsi(sphb200_group_last_error(g)) ?? "";
        if (rc == SPH_ERR_CAPACITY) throw new ArgumentOutOfRangeException("count", msg);
        if (rc == SPH_ERR_NEIGHBOR_OVERFLOW) throw new OverflowException(msg);
        if (rc == SPH_ERR_CUDA || rc == SPH_ERR_NCCL) throw new NotImplementedException("SPH group needs CUDA sm_100 devices and libnccl (no CPU fallback): " + msg);
        throw new InvalidOperationException("sphb200 group error " + rc + ": " + msg);
    }

    // Error mapping of SURVEY.md 8(b): the wrapper rethrows as the exception types the reference throws
    // (KernelSystem.cs:102-105 NotImplementedException, GravityFieldSystem.cs:400-403 InvalidOperationException).
    public static void Check(IntPtr h, int rc)
    {
        if (rc == SPH_OK) return;
        string msg = Marshal.PtrToStringAnsi(sphb200_last_error(h)) ?? "";
        switch (rc)
        {
            case SPH_ERR_CUDA: throw new NotImplementedException("SPH hot path needs a CUDA sm_100 device (no CPU fallback): " + msg);
            case SPH_ERR_STATE: throw new InvalidOperationException(msg);
            case SPH_ERR_CAPACITY: throw new ArgumentOutOfRangeException("count", msg);
            case SPH_ERR_NEIGHBOR_OVERFLOW: throw new OverflowException(msg);
            default: throw new InvalidOperationException("sphb200 error " + rc + ": " + msg);
        }
    }
}
