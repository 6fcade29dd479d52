// sph_systems.hpp -- C++ host mirror of the reference's six SPH systems over the libsphb200 C ABI.
//
// The reference's host side is compiled C# (Burst); with no C# toolchain in the build image this header is the compiled
// host layer above the C ABI: same system names, update order and constants as Assets/Scripts/Systems/*.cs, component
// arrays with the reference's byte layouts (include/sphb200.h).  One OnUpdate() = one native call.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/sphb200.h"

namespace sph {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("sphb200 error " + std::to_string(c) + ": " + m), code(c) {}
};

// ECS-style world: component arrays in body-index order + the device-resident simulation.
class World {
  public:
    std::vector<sph_Translation> Translation;
    std::vector<sph_PhysicsVelocity> PhysicsVelocity;
    std::vector<sph_ParticleMass> ParticleMass;
    std::vector<sph_ParticleSmoothing> ParticleSmoothing;
    std::vector<sph_ParticleDensity> ParticleDensity;
    std::vector<sph_ParticlePressure> ParticlePressure;
    std::vector<sph_ParticlePressureGrad> ParticlePressureGrad;
    std::vector<sph_GravityField> GravityField;
    float DeltaTime = 1.0f / 60.0f;
    sph_handle handle = nullptr;

    explicit World(int64_t count, int device = 0, const sph_Params* params = nullptr) {
        Translation.resize(count); PhysicsVelocity.resize(count); ParticleMass.resize(count); ParticleSmoothing.resize(count);
        ParticleDensity.resize(count); ParticlePressure.resize(count); ParticlePressureGrad.resize(count); GravityField.resize(count);
        int rc = sphb200_create(params, count, device, &handle);
        if (rc) throw Error(rc, sphb200_last_error(nullptr));
    }
    ~World() { if (handle) sphb200_destroy(handle); }
    World(const World&) = delete;
    World& operator=(const World&) = delete;

    void check(int rc) const { if (rc) throw Error(rc, sphb200_last_error(handle)); }
    int64_t count() const { return (int64_t)Translation.size(); }

    // BuildPhysicsWorld analogue (UP/ECS/Base/Systems/BuildPhysicsWorld.cs:389-469)
    void Upload() {
        check(sphb200_upload(handle, count(), Translation.data(), sizeof(sph_Translation), PhysicsVelocity.data(),
                             sizeof(sph_PhysicsVelocity), ParticleMass.data(), sizeof(sph_ParticleMass), ParticleSmoothing.data(),
                             sizeof(sph_ParticleSmoothing)));
    }
    // ExportPhysicsWorld analogue (UP/ECS/Base/Systems/ExportPhysicsWorld.cs:130-161) + SPH component write-back
    void Export() {
        check(sphb200_download(handle, SPH_FIELD_TRANSLATION, Translation.data(), sizeof(sph_Translation)));
        check(sphb200_download(handle, SPH_FIELD_VELOCITY, PhysicsVelocity.data(), sizeof(sph_PhysicsVelocity)));
        check(sphb200_download(handle, SPH_FIELD_SMOOTHING, ParticleSmoothing.data(), sizeof(sph_ParticleSmoothing)));
        check(sphb200_download(handle, SPH_FIELD_DENSITY, ParticleDensity.data(), sizeof(sph_ParticleDensity)));
        check(sphb200_download(handle, SPH_FIELD_PRESSURE, ParticlePressure.data(), sizeof(sph_ParticlePressure)));
        check(sphb200_download(handle, SPH_FIELD_PRESSURE_GRAD, ParticlePressureGrad.data(), sizeof(sph_ParticlePressureGrad)));
        check(sphb200_download(handle, SPH_FIELD_GRAVITY, GravityField.data(), sizeof(sph_GravityField)));
    }
};

// Several GPUs: the same component arrays over a GROUP handle (Morton-range domain decomposition, include/sphb200.h); one process
// drives every GPU.  The decomposition fuses the six stages into one call per fixed step (its exchanges sit between them).
class GroupWorld {
  public:
    std::vector<sph_Translation> Translation;
    std::vector<sph_PhysicsVelocity> PhysicsVelocity;
    std::vector<sph_ParticleMass> ParticleMass;
    std::vector<sph_ParticleSmoothing> ParticleSmoothing;
    std::vector<sph_ParticleDensity> ParticleDensity;
    std::vector<sph_GravityField> GravityField;
    float DeltaTime = 1.0f / 60.0f;
    sph_group group = nullptr;

    GroupWorld(int64_t count, const std::vector<int>& devices, const sph_Params* params = nullptr) {
        Translation.resize(count); PhysicsVelocity.resize(count); ParticleMass.resize(count); ParticleSmoothing.resize(count);
        ParticleDensity.resize(count); GravityField.resize(count);
        int rc = sphb200_group_create(params, count, (int)devices.size(), devices.data(), &group);
        if (rc) throw Error(rc, sphb200_group_last_error(nullptr));
    }
    ~GroupWorld() { if (group) sphb200_group_destroy(group); }
    GroupWorld(const GroupWorld&) = delete;
    GroupWorld& operator=(const GroupWorld&) = delete;
    void check(int rc) const { if (rc) throw Error(rc, sphb200_group_last_error(group)); }
    int64_t count() const { return (int64_t)Translation.size(); }
    void Upload() {
        check(sphb200_group_upload(group, count(), Translation.data(), sizeof(sph_Translation), PhysicsVelocity.data(), sizeof(sph_PhysicsVelocity),
                                   ParticleMass.data(), sizeof(sph_ParticleMass), ParticleSmoothing.data(), sizeof(sph_ParticleSmoothing)));
    }
    void Step(int gravity_impl) { check(sphb200_group_step(group, DeltaTime, gravity_impl)); }   // one FixedStepSimulationSystemGroup tick
    void Export() {
        check(sphb200_group_download(group, SPH_FIELD_TRANSLATION, Translation.data(), sizeof(sph_Translation)));
        check(sphb200_group_download(group, SPH_FIELD_VELOCITY, PhysicsVelocity.data(), sizeof(sph_PhysicsVelocity)));
        check(sphb200_group_download(group, SPH_FIELD_SMOOTHING, ParticleSmoothing.data(), sizeof(sph_ParticleSmoothing)));
        check(sphb200_group_download(group, SPH_FIELD_DENSITY, ParticleDensity.data(), sizeof(sph_ParticleDensity)));
        check(sphb200_group_download(group, SPH_FIELD_GRAVITY, GravityField.data(), sizeof(sph_GravityField)));
    }
};

struct SystemBase {
    World& world;
    explicit SystemBase(World& w) : world(w) {}
    virtual ~SystemBase() = default;
    virtual void OnUpdate() = 0;
};

struct ParticleSmoothingSystem : SystemBase {            // ParticleSmoothingSystem.cs:14-19
    static constexpr float TARGET_NEIGHBORS = 50.0f;
    using SystemBase::SystemBase;
    void OnUpdate() override { world.check(sphb200_smoothing_update(world.handle)); }
};
struct KernelSystem : SystemBase {                       // KernelSystem.cs:15-17
    using SystemBase::SystemBase;
    void OnUpdate() override { world.check(sphb200_build_neighbors(world.handle)); }
};
struct GravityFieldSystem : SystemBase {                 // GravityFieldSystem.cs:14-26, 228
    enum GravityImpl { GRAVITY_TREE_CPU = SPH_GRAVITY_TREE, GRAVITY_PARTICLE_CPU = SPH_GRAVITY_PARTICLE };
    static constexpr float k_GravConstant = 1.0f;
    static constexpr float k_Theta = 0.7f;
    GravityImpl k_GravityImpl = GRAVITY_TREE_CPU;
    using SystemBase::SystemBase;
    void OnUpdate() override { world.check(sphb200_gravity(world.handle, (int)k_GravityImpl, world.DeltaTime)); }
};
struct DensityFieldSystem : SystemBase {                 // DensityFieldSystem.cs:8-9
    using SystemBase::SystemBase;
    void OnUpdate() override { world.check(sphb200_density(world.handle)); }
};
struct PressureFieldSystem : SystemBase {                // PressureFieldSystem.cs:12-14
    using SystemBase::SystemBase;
    void OnUpdate() override { world.check(sphb200_pressure(world.handle)); }
};
struct VelocitySystem : SystemBase {                     // VelocitySystem.cs:15-16 (+ Integrator.cs:98-101)
    using SystemBase::SystemBase;
    void OnUpdate() override { world.check(sphb200_integrate(world.handle, world.DeltaTime)); }
};

// Update order of SURVEY.md section 3.1
struct FixedStepSimulationSystemGroup {
    ParticleSmoothingSystem smoothing;
    KernelSystem kernel;
    GravityFieldSystem gravity;
    DensityFieldSystem density;
    PressureFieldSystem pressure;
    VelocitySystem velocity;
    explicit FixedStepSimulationSystemGroup(World& w) : smoothing(w), kernel(w), gravity(w), density(w), pressure(w), velocity(w) {}
    void Update() {
        smoothing.OnUpdate(); kernel.OnUpdate(); gravity.OnUpdate(); density.OnUpdate(); pressure.OnUpdate(); velocity.OnUpdate();
    }
};

}  // namespace sph
