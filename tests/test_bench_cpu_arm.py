"""bench.py's CPU legs run without a GPU: the reference arm (`--impl reference`) and the `cpu_baseline` sampler are the
reference's job path restated in oracle/ and timed on the host cores (row d of SURVEY.md section 8).  These tests run them
on a scaled-down workload and check the contract of the JSON line the driver parses."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3",
                          "--particles", "30000", "--steps", "1", "--warmup", "0"], capture_output=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    lines = [l for l in out.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sph_particle_steps_per_sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "f32"
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["particles"] == 30000 and d["config"]["gravity"] == "particle"
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] == d["value"] and "EXTRAPOLATED" in c["sample"]
    det = c["detail"]
    # the reference's stage structure, stage by stage (KernelSystem.cs:93-231 and the systems after it)
    assert set(det["stage_us_per_particle"]) == {"smoothing", "aabb_bvh_build", "tree_overlap_candidates", "filter_pairs",
                                                 "flatten_pairs", "counting_sorts_x2", "calculate_interactions", "gravity",
                                                 "integrate_position", "density", "eos_pressure_gradient", "velocity_cleanup"}
    assert det["candidates_per_pair"] > 5.0                  # the broadphase hands ~15x the kept pairs to FilterPairs
    assert 30.0 < det["mean_neighbors"] < 80.0               # timed at the settled h (~50 own-support neighbors)
    assert det["cell_list_variant"]["label"].startswith("cell-list candidates") and det["cell_list_variant"]["particle_steps_per_sec"] > 0
    own = d["reference_own_cases"]                           # BASELINE.json configs[0] and [1], run in full
    assert own["c1"]["particles"] == 3000 and own["c1"]["gravity"] == "direct" and own["c1"]["particle_steps_per_sec"] > 0
    assert own["c2"]["particles"] == 10000 and own["c2"]["gravity"] == "tree" and own["c2"]["steps"] == 100
    # the reference's one published timing (README.md:33, KernelSystem stage at 3k particles) sits beside the port's time for that stage
    assert own["c1"]["kernel_system_stage_ms_published"] == 6.5
    assert 0.0 < own["c1"]["kernel_system_stage_ms_first"] < own["c1"]["ms_per_step_first"]
    assert 0.0 < own["c1"]["kernel_system_stage_ms_settled"] < own["c1"]["ms_per_step_settled"]


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.decode().strip() == ""


def test_cpu_sample_tree_workload_and_settled_h_estimate():
    import bench
    from sphb200 import ic
    c = ic.make_config("c4", particles=60000)
    n = len(c["h"])
    h = bench.settled_h_estimate(c)
    state = dict(pos=c["pos"], vel=c["vel"], mass=c["mass"], h=np.full(n, h, np.float32), n_own=np.full(n, 50, np.int32))
    v, cores, desc, det = bench.cpu_reference_sample(state, "tree", seconds_budget=1.0)
    assert v > 0 and cores >= 1 and "4-ary BVH" in desc and det["extrapolated"] is True
    # the estimate puts ~50 particles inside 2h: the sample's symmetric count lands near the controller's fixed point
    assert 35.0 < det["mean_neighbors"] < 70.0


def test_per_pass_bytes_add_up_to_the_survey_total():
    """368 + 12 K bytes per particle-step over the non-gravity passes; bench.py splits it by pass."""
    import bench
    K = 53.5
    ms = {"smoothing_bounds": 0.04, "keys_sort_permute_cells": 0.19, "neighbors_density_eos": 0.85, "pressure_grad": 0.29,
          "integrate": 0.018, "gravity_allpairs": 421.0}
    r = bench.per_pass_roofline(ms, K, 1 << 20, 6547.5)
    assert set(r) == set(ms) - {"gravity_allpairs"}
    total = sum(v["bytes_per_particle"] for k, v in r.items() if k != "smoothing_bounds")
    assert abs(total - (bench.SPH_BYTES_BASE + bench.SPH_BYTES_PER_NEIGHBOR * K)) < 1e-9
    v = r["integrate"]
    assert abs(v["achieved"] - 100.0 * (1 << 20) / 0.018e-3 / 1e9) < 1e-6 and abs(v["frac"] - v["achieved"] / 6547.5) < 1e-12
    assert bench.per_pass_roofline({"integrate": 0.0}, K, 10, 6547.5) == {}      # a pass that did not run is left out


def test_tree_walk_stats_from_the_reference_counters():
    import bench
    gf = np.zeros(1000, np.dtype([("value", "f4", 4), ("numParticles", "i4"), ("numApprox", "i4")]))
    gf["numParticles"] = 1200; gf["numApprox"] = 500
    r = bench.tree_walk_stats(gf, 2.0, 70.0)
    assert r["direct_per_particle"] == 1200.0 and r["approx_per_particle"] == 500.0 and r["interactions_per_particle"] == 1700.0
    assert abs(r["tflops_at_20_flop_per_interaction"] - 20.0 * 1700.0 * 1000 / 2.0e-3 / 1e12) < 1e-12
    assert abs(r["frac_of_fp32_peak"] - r["tflops_at_20_flop_per_interaction"] / 70.0) < 1e-15
    assert "tflops_at_20_flop_per_interaction" not in bench.tree_walk_stats(gf, 0.0, 70.0)
    assert bench.tree_walk_stats(gf[:0], 1.0, None)["interactions_per_particle"] == 0.0
