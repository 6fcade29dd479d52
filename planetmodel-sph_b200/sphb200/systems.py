"""Host-side mirror of the reference's six SPH systems (same names, order and constants), driving libsphb200.

Reference (all in Assets/Scripts/Systems/, group = FixedStepSimulationSystemGroup):
  ParticleSmoothingSystem  [UpdateBefore BuildPhysicsWorld]      ParticleSmoothingSystem.cs:14-19
  KernelSystem             [After Build, Before Step]            KernelSystem.cs:15-17
  GravityFieldSystem       [After Build, Before Step]            GravityFieldSystem.cs:14-16
  DensityFieldSystem       [After Step]                          DensityFieldSystem.cs:8-9
  PressureFieldSystem      [After Step, After Density]           PressureFieldSystem.cs:12-14
  VelocitySystem           [After ExportPhysicsWorld]            VelocitySystem.cs:15-16
plus Unity.Physics' integrator step (x += v dt) which here is folded into VelocitySystem's native call.

``World`` owns the ECS-style component arrays (byte-exact struct layouts); each ``OnUpdate`` completes its input
dependencies (a no-op: work is stream-ordered on the GPU) and makes exactly one native call.  The GPU state stays
resident between ticks; components are exported to the host arrays only by ``World.export()`` (the analogue of
ExportPhysicsWorld), so a tick costs no PCIe traffic unless the caller asks for data.
"""
import numpy as np
from . import (Simulation, Translation, PhysicsVelocity, ParticleSmoothing, GravityField, GRAVITY_TREE, GRAVITY_PARTICLE,
               FIELD_TRANSLATION, FIELD_VELOCITY, FIELD_SMOOTHING, FIELD_DENSITY, FIELD_PRESSURE, FIELD_PRESSURE_GRAD,
               FIELD_GRAVITY)


class JobHandle:
    """Unity's JobHandle stand-in: GPU work is ordered by the handle's CUDA stream, so every handle is 'default'."""

    def Complete(self):
        return None

    @staticmethod
    def CombineDependencies(a, b):
        return a if a is not None else b


class World:
    def __init__(self, count, device=0, dt=1.0 / 60.0, **params):
        self.count = int(count)
        self.DeltaTime = float(dt)
        self.Translation = np.zeros(count, Translation)
        self.PhysicsVelocity = np.zeros(count, PhysicsVelocity)
        self.ParticleMass = np.zeros(count, np.float32)
        self.ParticleSmoothing = np.zeros(count, ParticleSmoothing)
        self.ParticleDensity = np.zeros(count, np.float32)
        self.ParticlePressure = np.zeros(count, np.float32)
        self.ParticlePressureGrad = np.zeros((count, 3), np.float32)
        self.GravityField = np.zeros(count, GravityField)
        self.sim = Simulation(count, device=device, **params)
        self._systems = {}

    def set_particles(self, pos, vel, mass, h):
        """Authoring step (ParticleAuthoring.cs:150-245): fill the components, then BuildPhysicsWorld-style upload."""
        self.Translation["x"], self.Translation["y"], self.Translation["z"] = pos[:, 0], pos[:, 1], pos[:, 2]
        self.PhysicsVelocity["linear"] = vel
        self.ParticleMass[:] = mass
        self.ParticleSmoothing["influenceArea"] = h
        self.ParticleSmoothing["supportDomain"] = 2.0 * np.asarray(h, np.float32)  # ParticleSmoothing.cs:9-15
        self.ParticleSmoothing["neighbors"] = 0
        self.upload()

    def upload(self):
        self.sim.upload(self.Translation, self.PhysicsVelocity, self.ParticleMass, self.ParticleSmoothing)

    def export(self):
        """ExportPhysicsWorld analogue (UP/ECS/Base/Systems/ExportPhysicsWorld.cs:130-161) + component write-back."""
        s = self.sim
        s.download(FIELD_TRANSLATION, self.Translation)
        s.download(FIELD_VELOCITY, self.PhysicsVelocity)
        s.download(FIELD_SMOOTHING, self.ParticleSmoothing)
        s.download(FIELD_DENSITY, self.ParticleDensity)
        s.download(FIELD_PRESSURE, self.ParticlePressure)
        s.download(FIELD_PRESSURE_GRAD, self.ParticlePressureGrad)
        s.download(FIELD_GRAVITY, self.GravityField)

    def GetOrCreateSystem(self, cls):
        if cls not in self._systems:
            self._systems[cls] = cls(self)
        return self._systems[cls]

    GetExistingSystem = GetOrCreateSystem


class SystemBase:
    def __init__(self, world):
        self.World = world
        self.InputDependency = JobHandle()
        self.OutputDependency = JobHandle()

    # IPhysicsSystem (UP/ECS/Base/Systems/IPhysicsSystem.cs:6-11)
    def AddInputDependency(self, jh):
        self.InputDependency = JobHandle.CombineDependencies(jh, self.InputDependency)

    def GetOutputDependency(self):
        return self.OutputDependency

    def OnUpdate(self):
        raise NotImplementedError


class ParticleSmoothingSystem(SystemBase):
    TARGET_NEIGHBORS = 50.0  # ParticleSmoothingSystem.cs:18

    def OnUpdate(self):
        self.World.sim.smoothing_update()


class KernelSystem(SystemBase):
    def OnUpdate(self):
        self.InputDependency.Complete()
        self.World.sim.build_neighbors()


class GravityFieldSystem(SystemBase):
    GRAVITY_TREE_CPU = GRAVITY_TREE          # GravityFieldSystem.cs:19-23 (names kept; both run on the GPU here)
    GRAVITY_PARTICLE_CPU = GRAVITY_PARTICLE
    k_GravityImpl = GRAVITY_TREE             # :25
    k_GravConstant = 1.0                     # :26
    k_Theta = 0.7                            # :228

    def OnUpdate(self):
        self.InputDependency.Complete()
        self.World.sim.gravity(self.k_GravityImpl, self.World.DeltaTime)


class DensityFieldSystem(SystemBase):
    def OnUpdate(self):
        self.InputDependency.Complete()
        self.World.sim.density()


class PressureFieldSystem(SystemBase):
    def OnUpdate(self):
        self.World.sim.pressure()


class VelocitySystem(SystemBase):
    def OnUpdate(self):
        self.World.sim.integrate(self.World.DeltaTime)


class FixedStepSimulationSystemGroup:
    """Update order of SURVEY.md section 3.1."""
    ORDER = (ParticleSmoothingSystem, KernelSystem, GravityFieldSystem, DensityFieldSystem, PressureFieldSystem, VelocitySystem)

    def __init__(self, world, gravity_impl=None):
        self.world = world
        self.systems = [world.GetOrCreateSystem(c) for c in self.ORDER]
        if gravity_impl is not None:
            world.GetOrCreateSystem(GravityFieldSystem).k_GravityImpl = gravity_impl

    def Update(self):
        for s in self.systems:
            s.OnUpdate()
