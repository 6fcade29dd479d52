// SphB200Native.cs -- P/Invoke surface of libsphb200 (include/sphb200.h), the binding a PlanetModel-SPH maintainer
// adds to Assets/Scripts/.  Shipped as source: the build image has no C#/Unity toolchain, so this file is exercised
// only through its byte-layout contract (tests/test_abi_exports.py compiles the C header and checks every size/offset
// that the [StructLayout] attributes below assume).
using System;
using System.Runtime.InteropServices;

public static unsafe class SphB200Native
{
    const string Lib = "sphb200";   // libsphb200.so / sphb200.dll next to the player binary

    public const int SPH_OK = 0, SPH_ERR_INVALID_ARG = -1, SPH_ERR_CAPACITY = -2, SPH_ERR_NEIGHBOR_OVERFLOW = -3,
                     SPH_ERR_CUDA = -4, SPH_ERR_STATE = -5, SPH_ERR_TREE_STACK = -6;

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

    [DllImport(Lib)] public static extern int sphb200_default_params(Params* p);
    [DllImport(Lib)] public static extern int sphb200_create(Params* p, long capacity, int device, out IntPtr handle);
    [DllImport(Lib)] public static extern int sphb200_destroy(IntPtr h);
    [DllImport(Lib)] public static extern IntPtr sphb200_last_error(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_sync(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_upload(IntPtr h, long n, void* pos, int posStride, void* vel, int velStride,
                                                            void* mass, int massStride, void* smoothing, int smoothingStride);
    [DllImport(Lib)] public static extern int sphb200_smoothing_update(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_build_neighbors(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_gravity(IntPtr h, int impl, float dt);
    [DllImport(Lib)] public static extern int sphb200_density(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_pressure(IntPtr h);
    [DllImport(Lib)] public static extern int sphb200_integrate(IntPtr h, float dt);
    [DllImport(Lib)] public static extern int sphb200_step(IntPtr h, float dt, int gravityImpl);
    [DllImport(Lib)] public static extern int sphb200_download(IntPtr h, int field, void* dst, int stride);
    [DllImport(Lib)] public static extern int sphb200_download_neighbors(IntPtr h, long* offsets, int* nbr, long cap, long* total);
    [DllImport(Lib)] public static extern int sphb200_diagnostics(IntPtr h, double* out12);

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
