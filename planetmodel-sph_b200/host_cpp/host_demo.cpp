// host_demo.cpp -- C++ stand-in for the Unity player loop: spawns the reference scene (3000 particles, R = 50, M = 100,
// SimScene.unity:276-279) with a seeded LCG, ticks the six systems, prints conserved-quantity diagnostics.
//   ./host_demo [count] [steps] [tree|particle] [ranks]     ranks > 1: the same scene through a group handle (ranks 0..ranks-1 on
//   GPUs 0..ranks-1, or all on GPU 0 with SPH_DEMO_ONE_GPU=1: in-process transport)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "sph_systems.hpp"

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static double urand() {  // splitmix64 -> [0,1)
    uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char** argv) {
    int64_t count = argc > 1 ? atoll(argv[1]) : 3000;
    int steps = argc > 2 ? atoi(argv[2]) : 10;
    bool tree = !(argc > 3 && strcmp(argv[3], "particle") == 0);
    const float radius = 50.0f * (float)std::cbrt((double)count / 3000.0), totalMass = 100.0f * (float)count / 3000.0f;
    const float particleRadius = 5.0f;
    const int ranks = argc > 4 ? atoi(argv[4]) : 1;
    if (ranks > 1) {
        try {
            std::vector<int> devs(ranks);
            for (int r = 0; r < ranks; r++) devs[r] = getenv("SPH_DEMO_ONE_GPU") ? 0 : r;
            sph::GroupWorld world(count, devs);
            for (int64_t i = 0; i < count; i++) {
                float x, y, z;
                do {
                    x = (float)(2 * urand() - 1) * radius; y = (float)(2 * urand() - 1) * radius; z = (float)(2 * urand() - 1) * radius;
                } while (x * x + y * y + z * z > radius * radius);
                world.Translation[i] = {x, y, z};
                world.PhysicsVelocity[i] = {{0, 0, 0}, {0, 0, 0}};
                world.ParticleMass[i].value = totalMass / (float)count;
                float support = particleRadius * (1.0f + 0.5f * (float)urand());
                world.ParticleSmoothing[i] = {support / 2.0f, support, {0, 0, 0, 0}, 0};
            }
            world.Upload();
            for (int s = 0; s < steps; s++) {
                world.Step(tree ? SPH_GRAVITY_TREE : SPH_GRAVITY_PARTICLE);
                double d[12];
                world.check(sphb200_group_diagnostics(world.group, d));
                printf("step %3d  (%d ranks)  mass %.6g  |p| %.3e  E_kin %.6e  E_pot %.6e  E_int %.6e  neighbors mean %.1f max %d\n", s, ranks, d[0],
                       std::sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]), d[7], d[8], d[9], d[10], (int)d[11]);
            }
            world.Export();
            printf("particle 0: x = (%g, %g, %g)  rho = %g  gradPhi = (%g, %g, %g)\n", world.Translation[0].x, world.Translation[0].y,
                   world.Translation[0].z, world.ParticleDensity[0].value, world.GravityField[0].value[0], world.GravityField[0].value[1],
                   world.GravityField[0].value[2]);
        } catch (const sph::Error& e) {
            fprintf(stderr, "%s\n", e.what());
            return 1;
        }
        return 0;
    }
    try {
        sph::World world(count);
        for (int64_t i = 0; i < count; i++) {  // ParticleAuthoring.cs:229-245 rejection sampling, :208 equal masses
            float x, y, z;
            do {
                x = (float)(2 * urand() - 1) * radius; y = (float)(2 * urand() - 1) * radius; z = (float)(2 * urand() - 1) * radius;
            } while (x * x + y * y + z * z > radius * radius);
            world.Translation[i] = {x, y, z};
            world.PhysicsVelocity[i] = {{0, 0, 0}, {0, 0, 0}};
            world.ParticleMass[i].value = totalMass / (float)count;
            float support = particleRadius * (1.0f + 0.5f * (float)urand());
            world.ParticleSmoothing[i] = {support / 2.0f, support, {0, 0, 0, 0}, 0};
        }
        world.Upload();
        sph::FixedStepSimulationSystemGroup group(world);
        group.gravity.k_GravityImpl = tree ? sph::GravityFieldSystem::GRAVITY_TREE_CPU : sph::GravityFieldSystem::GRAVITY_PARTICLE_CPU;
        for (int s = 0; s < steps; s++) {
            group.Update();
            double d[12];
            world.check(sphb200_diagnostics(world.handle, d));
            printf("step %3d  mass %.6g  |p| %.3e  E_kin %.6e  E_pot %.6e  E_int %.6e  neighbors mean %.1f max %d\n", s, d[0],
                   std::sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]), d[7], d[8], d[9], d[10], (int)d[11]);
        }
        world.Export();
        printf("particle 0: x = (%g, %g, %g)  rho = %g  P = %g  gradPhi = (%g, %g, %g)\n", world.Translation[0].x, world.Translation[0].y,
               world.Translation[0].z, world.ParticleDensity[0].value, world.ParticlePressure[0].value, world.GravityField[0].value[0],
               world.GravityField[0].value[1], world.GravityField[0].value[2]);
    } catch (const sph::Error& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
