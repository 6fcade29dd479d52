#!/usr/bin/env python
"""bench.py -- SPH particle-steps/sec of the full hot path (smoothing + neighbors + density + EOS + pressure force +
gravity + integrate) on synthetic gas spheres, per BASELINE.json.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c2|c1] [--impl ours|reference]

Workloads (BASELINE.json configs): c3 = 1 048 576-particle sphere, tiled all-pairs gravity (default; the 1-GPU headline);
c4 = 16 000 000-particle sphere, LBVH tree gravity; c1/c2 = the reference's own 3k / 10k scenes.
One rank per GPU (torchrun for N > 1): Morton-range domain decomposition behind the C ABI (sphb200_group_*), halo exchange
and gravity-source all-gather over NCCL/NVLink inside the library.

Prints ONE JSON line (rank 0).  `value` = particles * K / max-over-ranks device time with the state resident in HBM;
`e2e` = the same through the C ABI with HOST component arrays uploaded and downloaded every step;
`roofline` = dominant kernel against its bound; `cpu_baseline` = the CPU oracle (port of the reference job path) on the
box's host cores over a bounded sample.  `--impl reference` times only that CPU port (the reference itself is C#/Unity
and cannot run here -- DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "planetmodel-sph_b200"))

import numpy as np  # noqa: E402

METRIC = "sph_particle_steps_per_sec"
UNIT = "particle-steps/s"
DT = 1.0 / 60.0
FLOP_PER_PAIR = 20.0          # conventional N-body count per ordered pair
SPH_BYTES_BASE, SPH_BYTES_PER_NEIGHBOR = 368.0, 12.0   # algorithmic HBM bytes / particle-step
NCU_TRAFFIC_ALLPAIRS_C3 = 136.4e6   # dram__bytes_read.sum + dram__bytes_write.sum of k_gravity_allpairs<6, equal-mass> at C3 (profiles/r02_allpairs_full.txt)


def workload_config(name, particles=None):
    from sphb200 import ic
    c = ic.make_config(name, particles=particles)
    grav = {"c1": "particle", "c2": "tree", "c3": "particle", "c4": "tree", "c5": "tree"}[name]
    return c, grav


# SURVEY.md 8(d) per pass (bytes per particle-step; K = mean symmetric neighbor count): the single handle's pass names.
# keys 24 + sort 68 + permute 68 + cell table 12; list emit 20 + 4K and density 44 + 4K; pressure force 32 + 4K; integrate 100.
# The h update (16 R + 4 W + 4 R counts) is DESIGN.md's own figure: SURVEY's 368 + 12K total does not count it.
PASS_BYTES = {"smoothing_bounds": (24.0, 0.0), "keys_sort_permute_cells": (172.0, 0.0), "neighbors_density_eos": (64.0, 8.0),
              "pressure_grad": (32.0, 4.0), "integrate": (100.0, 0.0)}


def per_pass_roofline(mean_ms, kbar, particles, hbm_peak_gbs):
    """Every non-gravity pass of the single handle against the HBM roof: algorithmic bytes / CUDA-event time of the pass."""
    out = {}
    for name, (base, per_k) in PASS_BYTES.items():
        ms = mean_ms.get(name)
        if not ms or ms <= 0:
            continue
        b = base + per_k * kbar
        gbs = b * particles / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "bytes_per_particle": b, "achieved": gbs, "frac": gbs / hbm_peak_gbs}
    return out


def tree_walk_stats(gf, walk_ms, fp32_peak_tflops):
    """SURVEY.md 8(d), tree gravity: interactions per particle from the reference's own counters (GravityField.numParticles =
    bodies summed directly, .numApprox = accepted node monopoles: GravityField.cs:14-15) x 20 flop, against the walk's time.
    No single roof is claimed for the walk; the FP32 fraction says how far the useful pair work sits below the FMA pipe."""
    n = len(gf)
    direct = float(gf["numParticles"].astype(np.float64).mean()) if n else 0.0
    approx = float(gf["numApprox"].astype(np.float64).mean()) if n else 0.0
    out = {"direct_per_particle": direct, "approx_per_particle": approx, "interactions_per_particle": direct + approx, "ms": walk_ms}
    if walk_ms and walk_ms > 0:
        out["tflops_at_20_flop_per_interaction"] = FLOP_PER_PAIR * (direct + approx) * n / (walk_ms * 1e-3) / 1e12
        if fp32_peak_tflops:
            out["frac_of_fp32_peak"] = out["tflops_at_20_flop_per_interaction"] / fp32_peak_tflops
    return out


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for k, nm in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(len(r) > k and r[k].lower().startswith("active") for r in self.rows):
                reasons.append(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "power_w_max": max(float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()) if self.rows else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def settled_h_estimate(c, target=50.0):
    """Smoothing length the reference's controller converges to on a sphere of this number density: 50 particles inside 2h
    (ParticleSmoothingSystem.cs:18).  The CPU arm cannot afford the ~10 steps the controller needs at 1M, so it starts there;
    the GPU arm's in-line baseline uses the state the GPU itself settled (same workload on both sides)."""
    n = len(c["h"])
    R = float(np.linalg.norm(c["pos"], axis=1).max())
    dens = n / (4.0 / 3.0 * np.pi * R ** 3)
    return float(0.5 * (3.0 * target / (4.0 * np.pi * dens)) ** (1.0 / 3.0))


def cpu_reference_sample(state, grav, seconds_budget=20.0, cell_variant=True):
    """Time the reference JOB PATH (oracle/sph_oracle.cpp orc_reference_step: Unity-shaped BVH build, dual-tree candidate
    pairs, FilterPairs, flatten, the two single-thread counting sorts, interaction buffers with both kernels evaluated on
    either side, density, EOS, pressure gradient, integration) on a bounded sample of the workload and extrapolate per
    particle: the SPH passes on a central ball of the state (same number density and h), gravity on a sample of targets
    against ALL sources.  `state` = dict(pos, vel, mass, h, n_own) with SETTLED smoothing lengths.
    Returns (particle_steps_per_sec, cores, description, detail dict)."""
    from oracle import oracle as orc
    n = len(state["h"])
    orc.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))   # torchrun exports OMP_NUM_THREADS=1: override for the CPU arm
    cores = orc.lib().orc_num_threads()
    # ~40 us of job path per particle and thread at ~50 neighbors (measured): size the ball for ~60 % of the budget
    ns = int(min(n, max(20000, 0.6 * seconds_budget * cores / 40e-6)))
    if ns < n:
        r = np.linalg.norm(state["pos"] - state["pos"].mean(0), axis=1)
        sel = np.argpartition(r, ns)[:ns]
        sel.sort()
    else:
        sel = np.arange(n)
    st = orc.State(state["pos"][sel], state["vel"][sel], state["mass"][sel], state["h"][sel], state["n_own"][sel])
    # n_own = 50 keeps h as it is in the sample step (radius ratio 1): the sample is timed AT the settled h
    st.n_own[:] = 50
    info = orc.reference_step(st, DT, gravity="none", want_lists=False)
    t_sph = sum(info["stage_sec"].values()) / ns
    kbar = info["interactions"] / ns
    cell = None
    if cell_variant:
        # -- the same SPH stages with CELL-LIST candidates (orc_neighbors_grid) in place of the reference's broadphase
        #    BVH + dual-tree overlap -- NOT the reference's path; printed beside it, labelled, never used for `value`
        c0 = time.perf_counter()
        off, nb = orc.neighbors(st.pos, st.h, "grid")
        c1 = time.perf_counter()
        rho_c, _ = orc.density(st.pos, st.h, st.mass, off, nb)
        P_c = orc.eos(rho_c)
        orc.pressure_grad(st.pos, st.h, st.mass, rho_c, P_c, off, nb)
        c2 = time.perf_counter()
        cell = {"label": "cell-list candidates instead of the BVH dual-tree overlap; one kernel evaluation per use "
                         "(not the reference's job path)",
                "neighbor_lists_us_per_particle": round(1e6 * (c1 - c0) / ns, 4),
                "density_eos_pressure_us_per_particle": round(1e6 * (c2 - c1) / ns, 4)}
        del off, nb, rho_c, P_c
    # -- gravity sample
    if grav == "particle":
        nt = int(max(64, min(n, 0.3 * seconds_budget * cores * 2.5e8 / n)))   # ~4 ns per pair and thread
        t0 = time.time()
        orc.gravity_direct(state["pos"], state["h"], state["mass"], i0=0, i1=nt)
        t_grav = (time.time() - t0) / nt
        gdesc = "direct gravity for %d targets x %d sources (x N/targets)" % (nt, n)
    else:
        ng = int(min(n, max(50000, 0.3 * seconds_budget * cores / 60e-6)))
        selg = sel[:ng] if ng <= len(sel) else np.arange(min(n, ng))
        sg = orc.State(state["pos"][selg], state["vel"][selg], state["mass"][selg], state["h"][selg], None)
        sg.n_own[:] = 50
        ig = orc.reference_step(sg, DT, gravity="tree", want_lists=False)
        t_grav = ig["stage_sec"]["gravity"] / len(selg) * (np.log(max(n, 2)) / np.log(max(len(selg), 2)))
        gdesc = "Unity-shaped 4-ary BVH moments + walk on %d particles (x log N / log n)" % len(selg)
    per_particle = t_sph + t_grav
    desc = ("EXTRAPOLATED from a bounded sample: reference job path (oracle C++ port, OpenMP; BVH dual-tree candidates x%.1f of "
            "the kept pairs, 2 serial counting sorts, 4 kernel evaluations per pair) on a central ball of %d of %d particles at "
            "settled h (%.1f neighbors avg), %s; per-particle times scaled to N=%d"
            % (info["candidates"] / max(info["pairs"], 1), ns, n, kbar, gdesc, n))
    detail = {"sample_particles": ns, "mean_neighbors": kbar, "candidates_per_pair": info["candidates"] / max(info["pairs"], 1),
              "stage_us_per_particle": {k: round(1e6 * v / ns, 4) for k, v in info["stage_sec"].items()},
              "gravity_us_per_particle": 1e6 * t_grav, "extrapolated": True, "cell_list_variant": cell}
    if cell is not None:
        cell["particle_steps_per_sec"] = 1.0 / ((c2 - c0) / ns + t_grav)
    return 1.0 / per_particle, cores, desc, detail


KERNEL_SYSTEM_STAGES = ("filter_pairs", "flatten_pairs", "counting_sorts_x2", "calculate_interactions")


def cpu_reference_own_cases():
    """The reference's own CPU-runnable cases IN FULL (BASELINE.json configs[0] and [1]): C1 = 3 000 particles, direct gravity;
    C2 = 10 000 particles, tree gravity, 100 steps.  Whole job path, every particle, wall clock per step."""
    from oracle import oracle as orc
    from sphb200 import ic
    orc.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))
    out = {}
    for name, grav, steps in (("c1", "direct", 20), ("c2", "tree", 100)):
        c = ic.make_config(name)
        st = orc.State(c["pos"], c["vel"], c["mass"], c["h"])
        t, ks = [], []
        for _ in range(steps):
            t0 = time.perf_counter()
            info = orc.reference_step(st, DT, gravity=grav, want_lists=False)
            t.append(time.perf_counter() - t0)
            ks.append(sum(info["stage_sec"][k] for k in KERNEL_SYSTEM_STAGES))
        half = max(steps // 2, 1)
        last = float(np.mean(t[-half:]))        # settled half of the run
        out[name] = {"particles": len(c["h"]), "gravity": grav, "steps": steps, "ms_per_step_settled": 1e3 * last,
                     "ms_per_step_first": 1e3 * t[0], "particle_steps_per_sec": len(c["h"]) / last,
                     "mean_neighbors_final": info["interactions"] / len(c["h"]), "kind": "port, run in full",
                     # KernelSystem.OnUpdate alone (FilterPairs, flatten, the two counting sorts, CalculateInteraction)
                     "kernel_system_stage_ms_first": 1e3 * ks[0], "kernel_system_stage_ms_settled": 1e3 * float(np.mean(ks[-half:]))}
    # the one timing the reference publishes: that stage at 3 000 particles on the author's laptop
    out["c1"]["kernel_system_stage_ms_published"] = 6.5
    out["c1"]["published_source"] = "reference README.md:33 (unspecified gaming laptop, Unity + Burst); sanity anchor, not a baseline"
    return out


def run_reference(args):
    """--impl reference: the reference's CPU job path (oracle port: the C# cannot be built here) on the host cores; every step
    is a bounded sample of the workload, extrapolated per particle (stated in the line)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args.workload = args.workload or "c3"
    c, grav = workload_config(args.workload, args.particles)
    n = len(c["h"])
    state = dict(pos=c["pos"], vel=c["vel"], mass=c["mass"], h=np.full(n, settled_h_estimate(c), np.float32),
                 n_own=np.full(n, 50, np.int32))
    budget = max(1.0, min(20.0, 100.0 / max(args.warmup + args.steps, 1)))     # nominal seconds per sampled step: the whole run ends within a few minutes (a host half as fast as the B200 boxes takes twice the nominal time)
    vals, wall = [], []
    for k in range(args.warmup + args.steps):
        t0 = time.time()
        # the labelled cell-list variant rides on the last step only (it is not part of the timed job path)
        v, cores, desc, detail = cpu_reference_sample(state, grav, budget, cell_variant=(k == args.warmup + args.steps - 1))
        if k >= args.warmup:
            vals.append(v); wall.append(time.time() - t0)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * n / v, "ms_per_step_is": "extrapolated from the sample (not wall clock)",
            "sample_wall_s_per_step": float(np.mean(wall)), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_desc(args.workload, n, grav, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc, "detail": detail},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "reference_own_cases": cpu_reference_own_cases()}
    emit(line)


def workload_desc(name, n, grav, gpus):
    return {"workload": "%s: %d-particle %s, %s gravity, dt=1/60, full step" %
            (name, n, "two-planet collision (64x density contrast)" if name == "c5" else "uniform gas sphere (reference scene density)",
             "tiled all-pairs" if grav == "particle" else "LBVH Barnes-Hut theta=0.7"),
            "particles": n, "gravity": grav, "parallelism": "morton-range decomposition x%d" % gpus,
            "l2_policy": "working set (>= 190 MB of SoA + lists at 1M) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    """Default: the headline workload C3 and -- in the same invocation and JSON line, under the key "c4" -- the 16 M
    tree-gravity workload C4 (BASELINE.json's metric is quoted at 1M and at 16M).  --workload X measures only X."""
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    names = [args.workload] if args.workload else ["c3", "c4"]
    line = None
    for k, name in enumerate(names):
        rec = measure_workload(args, name, world, rank, local, headline=(k == 0))
        if rank == 0:
            if k == 0:
                line = rec
            else:
                line[name] = rec
    if rank == 0:
        emit(line)
    if world > 1:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()


class Engine:
    """N = 1: one handle (sphb200_*).  N > 1: one rank of a Morton-range-decomposed group (sphb200_group_*), one process per
    GPU, NCCL over NVLink inside the library.  Either way every process holds only its body slice of the host state."""

    def __init__(self, c, world, rank, local, **params):
        import sphb200
        from sphb200 import group as sgroup
        self.world, self.rank = world, rank
        self.n = len(c["h"])
        if world == 1:
            self.sim = sphb200.Simulation(self.n, device=local, **params)
            self.b0, self.cnt = 0, self.n
        else:
            self.sim = sgroup.Group.from_env(self.n, device=local, **params)
            self.b0, self.cnt = self.sim.body_range(self.n)
        self.slice = {k: np.ascontiguousarray(v[self.b0:self.b0 + self.cnt]) for k, v in c.items()}

    def upload(self, pos, vel, mass, sm):
        if self.world == 1:
            self.sim.upload(pos, vel, mass, sm)
        else:
            self.sim.upload(self.n, pos, vel, mass, sm)

    def step(self, dt, impl):
        self.sim.step(dt, impl)

    def close(self):
        self.sim.close()


def measure_workload(args, workload, world, rank, local, headline):
    import torch
    import sphb200
    c, grav = workload_config(workload, args.particles)
    if args.gravity:
        grav = args.gravity
    n = len(c["h"])
    impl = sphb200.GRAVITY_PARTICLE if grav == "particle" else sphb200.GRAVITY_TREE

    extra = {}
    if args.leaf_max:
        extra["leaf_max"] = args.leaf_max
    if args.aabb_mode:
        extra["aabb_mode"] = args.aabb_mode
    eng = Engine(c, world, rank, local, **extra)
    sl = eng.slice
    eng.upload(sl["pos"], sl["vel"], sl["mass"], sl["h"])
    sim = eng.sim

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also settles h towards ~50 own-support neighbors)
    for _ in range(args.warmup):
        eng.step(DT, impl)
    sim.sync()
    barrier()
    sim.enable_timing(True)
    launches0 = sim.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    stream = torch.cuda.ExternalStream(sim.stream_ptr(), device=torch.device("cuda", local))   # the library's own stream
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    pass_ms = {}
    e0.record(stream)
    for _ in range(args.steps):
        eng.step(DT, impl)
        for nm, ms in sim.timings():            # reads the CUDA events of the step just issued (all ranks alike)
            pass_ms.setdefault(nm, []).append(ms)
    e1.record(stream)
    e1.synchronize()
    barrier()
    sampler.stop_flag = True
    ms_total = e0.elapsed_time(e1)
    launches = sim.launch_count() - launches0
    if world > 1:
        import torch.distributed as td
        t = torch.tensor([ms_total], device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_total = float(t.item())
        t = torch.tensor([float(launches)], device="cuda")
        td.all_reduce(t)
        launches = int(t.item())
    ms_per_step = ms_total / args.steps
    value = n * args.steps / (ms_total * 1e-3)
    sim.enable_timing(False)
    errors = None
    try:
        sim.sync()                      # surfaces sticky asynchronous errors (neighbor-list overflow, tree stack)
    except sphb200.SphError as ex:
        errors = str(ex)
    info = sim.info() if world > 1 else None
    rank_ms = None
    if world > 1:      # per-rank times of the heaviest passes (load balance of the decomposition)
        import torch.distributed as td
        names = sorted(pass_ms)
        mine = torch.tensor([float(np.mean(pass_ms[k])) for k in names], device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        td.all_gather(allr, mine)
        rank_ms = {k: [round(float(a[i]), 3) for a in allr] for i, k in enumerate(names)}
    if rank == 0:
        print("[bench] %s: %d steps, %.3f ms/step, passes %s%s" % (workload, args.steps, ms_per_step,
              {k: round(float(np.mean(v)), 3) for k, v in pass_ms.items()},
              "" if info is None else ", rank 0 owns %s + halo %s, %d migrated last step" % (info["n_own"], info["n_halo"], info["migrated_last_step"])),
              file=sys.stderr)

    diag = sim.diagnostics()            # collective in a group; the state the timed steps ran on (settled h)
    gf_last = None
    if grav == "tree" and world == 1:   # counters of the last timed step (single handle: no collective involved)
        try:
            gf_last = sim.download(sphb200.FIELD_GRAVITY, allow_overflow=True)
        except Exception:               # a reporting extra must never cost the line
            gf_last = None
    # ---- the settled state as the host holds it (every process: its body slice): input of the e2e leg and of the CPU baseline
    state = None
    if not args.kernels_only:
        smr = sim.download(sphb200.FIELD_SMOOTHING, allow_overflow=True)
        state = dict(pos=sim.download(sphb200.FIELD_TRANSLATION, allow_overflow=True), vel=sim.download(sphb200.FIELD_VELOCITY, allow_overflow=True),
                     mass=sim.download(sphb200.FIELD_MASS, allow_overflow=True), sm=smr)
    # ---- e2e: host component arrays in, host component arrays out, every step
    e2e = None
    if not args.kernels_only:
        e2e = measure_e2e(eng, state, impl, max(1, min(args.steps, 3)), barrier)
    if rank != 0:
        eng.close()
        return None
    # ---- roofline of the dominant kernel
    kbar = diag["mean_neighbors"]
    mean = {k: float(np.mean(v)) for k, v in pass_ms.items()}
    fp32_peak = None
    if world == 1:
        fp32_peak = sim.fp32_peak_tflops()
    else:
        probe = sphb200.Simulation(1024, device=local)
        fp32_peak = probe.fp32_peak_tflops()
        probe.close()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    # every pass that is not gravity: bounds, keys/migration/sort, halo exchange + cell table, neighbors + density, pressure, integrate
    sph_ms = sum(v for k, v in mean.items() if not k.startswith("gravity"))
    sph_bytes = (SPH_BYTES_BASE + SPH_BYTES_PER_NEIGHBOR * kbar) * n / world
    hbm_passes = {"achieved": sph_bytes / (sph_ms * 1e-3) / 1e9 if sph_ms > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                  "frac": (sph_bytes / (sph_ms * 1e-3) / 1e9 / hbm_peak) if sph_ms > 0 else None, "ms": sph_ms,
                  "bytes_per_particle": SPH_BYTES_BASE + SPH_BYTES_PER_NEIGHBOR * kbar, "peak_source": hbm_src}
    if world == 1:
        try:
            hbm_passes["per_pass"] = per_pass_roofline(mean, float(kbar), n, hbm_peak)
        except Exception as ex:      # a reporting extra must never cost the line
            hbm_passes["per_pass"] = {"error": str(ex)}
    if grav == "particle":
        gms = mean.get("gravity_allpairs", 0.0)
        flops = FLOP_PER_PAIR * (n / world) * (n - 1)
        ach = flops / (gms * 1e-3) / 1e12 if gms > 0 else None
        roof = {"bound": "fp32", "kernel": "k_gravity_allpairs", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach / fp32_peak if ach else None,
                # DRAM bytes per launch of this kernel at C3 / 1 GPU: dram__bytes_read.sum + dram__bytes_write.sum of the committed
                # ncu --set full capture (profiles/); ncu cannot run inside a timed bench, so it is not re-measured here
                "traffic": NCU_TRAFFIC_ALLPAIRS_C3 if (world == 1 and n == (1 << 20)) else None,
                "note": "FP32 FMA-pipe bound (not a contraction: no tensor roof applies); peak = FMA microbenchmark measured "
                        "in this run; flops = 20 per ordered pair (SURVEY 8d); the pass also holds the source all-gather at N > 1",
                "ms": gms, "share_of_step": gms / ms_per_step}
    else:
        gms = mean.get("gravity_tree", 0.0)
        roof = {"bound": "hbm", "kernel": "sph passes (keys/sort/cells/neighbors+density/pressure/integrate)",
                "achieved": hbm_passes["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": hbm_passes["frac"], "traffic": None,
                "note": "tree walk is L2-latency/FP32 mixed (no single roof, SURVEY 8d): %.2f ms of the step" % gms, "ms": sph_ms,
                "share_of_step": sph_ms / ms_per_step}
        if gf_last is not None:
            try:
                roof["tree_walk"] = tree_walk_stats(gf_last, gms, fp32_peak)
            except Exception as ex:
                roof["tree_walk"] = {"error": str(ex)}
    cpu_detail = None
    if args.kernels_only or world > 1 or not headline:
        cpu_v, cores, desc = None, 0, "skipped (%s)" % ("--kernels-only profiling run" if args.kernels_only else
                                                        "reported at N=1 for the headline workload only; see --impl reference")
    else:
        # the very state the GPU arm was timed on (settled h), as the host holds it
        cst = dict(pos=state["pos"], vel=state["vel"], mass=state["mass"], h=state["sm"]["influenceArea"].copy(),
                   n_own=state["sm"]["neighbors"].copy())
        cpu_v, cores, desc, cpu_detail = cpu_reference_sample(cst, grav)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_desc(workload, n, grav, world), "clocks": sampler.summary(),
            "e2e": e2e, "gpu_launches": int(launches), "errors": errors, "roofline": roof, "hbm_passes": hbm_passes,
            "pass_ms": mean, "mean_neighbors": kbar, "fp32_peak_tflops_measured": fp32_peak,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc, "detail": cpu_detail}}
    if cpu_detail is not None:
        line["cpu_baseline"]["reference_own_cases"] = cpu_reference_own_cases()
    if rank_ms is not None:
        line["pass_ms_per_rank"] = {k: v for k, v in rank_ms.items() if max(v) > 0.5}
    if info is not None:
        line["decomposition"] = {"rank0_own": info["n_own"], "rank0_halo": info["n_halo"], "migrated_last_step": info["migrated_last_step"],
                                 "tree_nodes_received_last_step": info.get("tree_nodes_last_step"), "transport": info["transport"]}
    eng.close()
    return line


def measure_e2e(eng, state, impl, steps, barrier):
    """Drop-in usage: ECS owns the components on the host; every step uploads them (pinned) and reads all results back.
    In a group every process moves only its body slice (1/N of the state) over PCIe; the particles travel between the
    GPUs over NVLink inside the step.  Wall clock between two barriers."""
    import torch
    import sphb200
    sim = eng.sim
    n, cnt = eng.n, eng.cnt
    # host component arrays (pinned), holding the settled state the timed steps ended on
    host = {
        "pos": torch.from_numpy(state["pos"].copy().reshape(-1)).pin_memory(),
        "vel": torch.from_numpy(state["vel"].copy().reshape(-1)).pin_memory(),
        "mass": torch.from_numpy(state["mass"].copy()).pin_memory(),
        "sm": torch.from_numpy(np.zeros(cnt * 7, np.float32)).pin_memory(),
    }
    sm = host["sm"].numpy().view(sphb200.ParticleSmoothing)
    sm[:] = state["sm"]
    outs = {f: torch.empty(cnt * w, dtype=torch.float32).pin_memory() for f, w in
            ((sphb200.FIELD_DENSITY, 1), (sphb200.FIELD_PRESSURE, 1), (sphb200.FIELD_PRESSURE_GRAD, 3), (sphb200.FIELD_GRAVITY, 6))}
    # Translation / PhysicsVelocity / ParticleSmoothing are read AND written by the step (ExportPhysicsWorld writes the components
    # in place): the download lands in the very arrays the next step uploads
    outs[sphb200.FIELD_TRANSLATION] = host["pos"]; outs[sphb200.FIELD_VELOCITY] = host["vel"]
    h2d = n * (12 + 12 + 4 + 28)                  # Translation, PhysicsVelocity.linear, ParticleMass, ParticleSmoothing records
    d2h = n * (12 + 12 + 4 + 4 + 12 + 24 + 28)    # ... + density, pressure, pressure gradient, GravityField, ParticleSmoothing

    phases = {"upload": 0.0, "step": 0.0, "download": 0.0}

    def one():
        ta = time.perf_counter()
        eng.upload(host["pos"].numpy().reshape(cnt, 3), host["vel"].numpy().reshape(cnt, 3), host["mass"].numpy(), sm)
        tb = time.perf_counter()
        eng.step(DT, impl)
        tc = time.perf_counter()                  # asynchronous: the step's device time shows up in the first download
        # the SPH fields first: they are final behind the pressure pass and their copies run beside the gravity pass
        # (single handle: sphb200_download moves them on the auxiliary stream); then what the end of the step produces
        def pull(f):
            buf = outs[f]
            w = buf.numel() // max(cnt, 1)
            sim.download(f, buf.numpy().reshape(cnt, w) if w > 1 else buf.numpy(), allow_overflow=True)
        for f in (sphb200.FIELD_DENSITY, sphb200.FIELD_PRESSURE, sphb200.FIELD_PRESSURE_GRAD):
            pull(f)
        sim.download(sphb200.FIELD_SMOOTHING, sm, allow_overflow=True)
        for f in (sphb200.FIELD_GRAVITY, sphb200.FIELD_TRANSLATION, sphb200.FIELD_VELOCITY):
            pull(f)
        td_ = time.perf_counter()
        phases["upload"] += tb - ta; phases["step"] += tc - tb; phases["download"] += td_ - tc
    one()
    for k in phases:
        phases[k] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    barrier()
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "host_phase_ms": {k: round(1e3 * v / steps, 3) for k, v in phases.items()},
            "path": "sphb200%s_upload + _step + _download x7 per process; every process moves its body slice (1/%d of the bytes) "
                    "between pinned host component arrays and its GPU, one DMA per component array" % ("_group" if eng.world > 1 else "", eng.world)}


_REAL_STDOUT = None


def emit(line):
    """Write the ONE JSON line to the real stdout (everything else, e.g. NCCL's banner, was diverted to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    # libraries (NCCL) print banners on fd 1: keep stdout for the single JSON line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["c1", "c2", "c3", "c4", "c5"],
                    help="measure only this workload (default: c3 as the headline plus c4 as a sub-record)")
    ap.add_argument("--particles", type=int, default=None, help="scale c3/c4/c5 down (or up) to this many particles")
    ap.add_argument("--gravity", default=None, choices=["tree", "particle"], help="override the workload's gravity path")
    ap.add_argument("--leaf-max", type=int, default=0, help="bodies per tree leaf (reference: 4)")
    ap.add_argument("--aabb-mode", type=int, default=0, help="1 = point-bounds MAC boxes (non-reference; quirk Q2 off)")
    ap.add_argument("--kernels-only", action="store_true", help="skip the e2e and CPU-baseline legs (ncu profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
