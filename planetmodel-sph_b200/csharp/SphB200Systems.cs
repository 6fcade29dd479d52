// SphB200Systems.cs -- drop-in replacements of the six SPH systems (Assets/Scripts/Systems/*.cs): same class names,
// same [UpdateInGroup]/[UpdateBefore]/[UpdateAfter] ordering, same IPhysicsSystem surface and public constants; each
// OnUpdate completes its dependencies and makes exactly one native call.  Particles stop being Unity.Physics rigid
// bodies (README.md:82-93 roadmap): no PostBroadphase callback, no PhysicsCollider, no BuildPhysicsWorld round trip.
//
// Data flow per fixed step:
//   SphWorldBridge.Upload   (once, or whenever authoring changes particles)   ECS chunks -> device SoA
//   six systems             (every step)                                       device only, no PCIe traffic
//   SphWorldBridge.Export   (when rendering / other systems need the data)     device SoA -> ECS components
// Source only (no C# toolchain in the build image); the C ABI it binds is exercised by the Python and C++ mirrors.
using System;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using Unity.Entities;
using Unity.Jobs;
using Unity.Mathematics;
using Unity.Physics;
using Unity.Physics.Systems;
using Unity.Transforms;

public unsafe class SphWorldBridge : IDisposable
{
    public IntPtr Handle;
    public int Count;
    public static SphWorldBridge Instance;

    public SphWorldBridge(int capacity, int device = 0)
    {
        SphB200Native.Params p;
        SphB200Native.sphb200_default_params(&p);
        p.K = 1000.0f;                                              // PressureFieldSystem.cs:31
        p.G = GravityFieldSystem.k_GravConstant;                    // GravityFieldSystem.cs:26
        p.theta = GravityFieldSystem.k_Theta;                       // GravityFieldSystem.cs:228
        p.target_neighbors = ParticleSmoothingSystem.TARGET_NEIGHBORS;
        SphB200Native.Check(IntPtr.Zero, SphB200Native.sphb200_create(&p, capacity, device, out Handle));
    }

    // Entity order of the query = body index (BuildPhysicsWorld.cs:389-469 used chunk iteration order the same way).
    public void Upload(EntityQuery q)
    {
        using (var pos = q.ToComponentDataArray<Translation>(Allocator.TempJob))
        using (var vel = q.ToComponentDataArray<PhysicsVelocity>(Allocator.TempJob))
        using (var mass = q.ToComponentDataArray<ParticleMass>(Allocator.TempJob))
        using (var sm = q.ToComponentDataArray<ParticleSmoothing>(Allocator.TempJob))
        {
            Count = pos.Length;
            SphB200Native.Check(Handle, SphB200Native.sphb200_upload(Handle, Count,
                pos.GetUnsafeReadOnlyPtr(), sizeof(Translation), vel.GetUnsafeReadOnlyPtr(), sizeof(PhysicsVelocity),
                mass.GetUnsafeReadOnlyPtr(), sizeof(ParticleMass), sm.GetUnsafeReadOnlyPtr(), sizeof(ParticleSmoothing)));
        }
    }

    // ExportPhysicsWorld.cs:130-161 analogue + write-back of the SPH components.
    public void Export(EntityQuery q)
    {
        void Pull<T>(SphB200Native.Field f) where T : struct, IComponentData
        {
            using (var a = new NativeArray<T>(Count, Allocator.TempJob))
            {
                SphB200Native.Check(Handle, SphB200Native.sphb200_download(Handle, (int)f, a.GetUnsafePtr(), UnsafeUtility.SizeOf<T>()));
                q.CopyFromComponentDataArray(a);
            }
        }
        Pull<Translation>(SphB200Native.Field.Translation);
        Pull<PhysicsVelocity>(SphB200Native.Field.Velocity);        // linear written, angular left zero
        Pull<ParticleSmoothing>(SphB200Native.Field.Smoothing);
        Pull<ParticleDensity>(SphB200Native.Field.Density);
        Pull<ParticlePressure>(SphB200Native.Field.Pressure);
        Pull<ParticlePressureGrad>(SphB200Native.Field.PressureGrad);
        Pull<GravityField>(SphB200Native.Field.Gravity);
    }

    public void Dispose() { if (Handle != IntPtr.Zero) { SphB200Native.sphb200_destroy(Handle); Handle = IntPtr.Zero; } }
}

// Several GPUs (the reference caps a world at 2^24-2 bodies, UP/Dynamics/Simulation/Scheduler.cs:24-41; a group does not): the same
// bridge over a group handle.  The decomposition fuses the six stages into one call per fixed step (its exchanges sit between
// them), so with a group the six systems stay registered for their ordering attributes and constants but only VelocitySystem --
// the last of the step -- makes the native call: SphGroupBridge.Instance != null switches them (see SphSystemBase.Stage).
public unsafe class SphGroupBridge : IDisposable
{
    public IntPtr Group;
    public int Count;
    public static SphGroupBridge Instance;

    public SphGroupBridge(int capacity, int[] devices)
    {
        SphB200Native.Params p;
        SphB200Native.sphb200_default_params(&p);
        p.G = GravityFieldSystem.k_GravConstant; p.theta = GravityFieldSystem.k_Theta; p.target_neighbors = ParticleSmoothingSystem.TARGET_NEIGHBORS;
        fixed (int* d = devices)
            SphB200Native.CheckGroup(IntPtr.Zero, SphB200Native.sphb200_group_create(&p, capacity, devices.Length, d, out Group));
    }

    public void Upload(EntityQuery q)
    {
        using (var pos = q.ToComponentDataArray<Translation>(Allocator.TempJob))
        using (var vel = q.ToComponentDataArray<PhysicsVelocity>(Allocator.TempJob))
        using (var mass = q.ToComponentDataArray<ParticleMass>(Allocator.TempJob))
        using (var sm = q.ToComponentDataArray<ParticleSmoothing>(Allocator.TempJob))
        {
            Count = pos.Length;
            SphB200Native.CheckGroup(Group, SphB200Native.sphb200_group_upload(Group, Count,
                pos.GetUnsafeReadOnlyPtr(), sizeof(Translation), vel.GetUnsafeReadOnlyPtr(), sizeof(PhysicsVelocity),
                mass.GetUnsafeReadOnlyPtr(), sizeof(ParticleMass), sm.GetUnsafeReadOnlyPtr(), sizeof(ParticleSmoothing)));
        }
    }

    public void Step(float dt, int impl) { SphB200Native.CheckGroup(Group, SphB200Native.sphb200_group_step(Group, dt, impl)); }

    public void Export<T>(EntityQuery q, SphB200Native.Field f) where T : struct, IComponentData
    {
        using (var a = new NativeArray<T>(Count, Allocator.TempJob))
        {
            SphB200Native.CheckGroup(Group, SphB200Native.sphb200_group_download(Group, (int)f, a.GetUnsafePtr(), UnsafeUtility.SizeOf<T>()));
            q.CopyFromComponentDataArray(a);
        }
    }

    public void Dispose() { if (Group != IntPtr.Zero) { SphB200Native.sphb200_group_destroy(Group); Group = IntPtr.Zero; } }
}

public abstract class SphSystemBase : SystemBase, IPhysicsSystem
{
    // IPhysicsSystem (UP/ECS/Base/Systems/IPhysicsSystem.cs:6-11): GPU work is stream-ordered, handles stay default.
    protected JobHandle InputDependency;
    public JobHandle GetOutputDependency() => default;
    public void AddInputDependency(JobHandle jh) => InputDependency = JobHandle.CombineDependencies(jh, InputDependency);
    protected IntPtr H => SphWorldBridge.Instance.Handle;
    protected void Begin() { InputDependency.Complete(); InputDependency = default; }
    // with a group the whole step is one native call, made by the last system of the step
    protected static bool Grouped => SphGroupBridge.Instance != null;
}

[UpdateInGroup(typeof(FixedStepSimulationSystemGroup))]
public class ParticleSmoothingSystem : SphSystemBase
{
    public const float TARGET_NEIGHBORS = 50;                        // ParticleSmoothingSystem.cs:18
    protected override void OnUpdate() { Begin(); if (Grouped) return; SphB200Native.Check(H, SphB200Native.sphb200_smoothing_update(H)); }
}

[UpdateInGroup(typeof(FixedStepSimulationSystemGroup))]
[UpdateAfter(typeof(ParticleSmoothingSystem))]
public class KernelSystem : SphSystemBase
{
    protected override void OnUpdate() { Begin(); if (Grouped) return; SphB200Native.Check(H, SphB200Native.sphb200_build_neighbors(H)); }
}

[UpdateInGroup(typeof(FixedStepSimulationSystemGroup))]
[UpdateAfter(typeof(KernelSystem))]
public class GravityFieldSystem : SphSystemBase
{
    public enum GravityImpl : ushort { GRAVITY_TREE_CPU, GRAVITY_PARTICLE_CPU }   // names kept; both run on the GPU
    public const GravityImpl k_GravityImpl = GravityImpl.GRAVITY_TREE_CPU;        // GravityFieldSystem.cs:25
    public const float k_GravConstant = 1.0f;                                     // :26
    public const float k_Theta = 0.7f;                                            // :228
    protected override void OnUpdate()
    {
        Begin();
        if (Grouped) return;
        int impl = k_GravityImpl == GravityImpl.GRAVITY_TREE_CPU ? SphB200Native.SPH_GRAVITY_TREE : SphB200Native.SPH_GRAVITY_PARTICLE;
        SphB200Native.Check(H, SphB200Native.sphb200_gravity(H, impl, World.Time.DeltaTime));
    }
}

[UpdateInGroup(typeof(FixedStepSimulationSystemGroup))]
[UpdateAfter(typeof(GravityFieldSystem))]
public class DensityFieldSystem : SphSystemBase
{
    protected override void OnUpdate() { Begin(); if (Grouped) return; SphB200Native.Check(H, SphB200Native.sphb200_density(H)); }
}

[UpdateInGroup(typeof(FixedStepSimulationSystemGroup))]
[UpdateAfter(typeof(DensityFieldSystem))]
public class PressureFieldSystem : SphSystemBase
{
    protected override void OnUpdate() { Begin(); if (Grouped) return; SphB200Native.Check(H, SphB200Native.sphb200_pressure(H)); }
}

[UpdateInGroup(typeof(FixedStepSimulationSystemGroup))]
[UpdateAfter(typeof(PressureFieldSystem))]
public class VelocitySystem : SphSystemBase
{
    // x += v dt (Integrator.cs:98-101) and v += (-gradP/rho - gradPhi) dt (VelocitySystem.cs:24-36) in one kernel
    protected override void OnUpdate()
    {
        Begin();
        if (Grouped)
        {
            int impl = GravityFieldSystem.k_GravityImpl == GravityFieldSystem.GravityImpl.GRAVITY_TREE_CPU ? SphB200Native.SPH_GRAVITY_TREE : SphB200Native.SPH_GRAVITY_PARTICLE;
            SphGroupBridge.Instance.Step(World.Time.DeltaTime, impl);     // all six stages + the exchanges between them, on every GPU
            return;
        }
        SphB200Native.Check(H, SphB200Native.sphb200_integrate(H, World.Time.DeltaTime));
    }
}
