"""bench.py's GPU arm cannot run here (no device), but everything around the device calls can: this test drives
bench.measure_workload with a stand-in engine (synthetic arrays, fixed pass times; NO physics and no oracle behind it) and a
stubbed torch.cuda, and checks that the ONE JSON line comes out complete, serialisable and with the keys the driver reads.
It guards the host-side assembly code (roofline, per-pass split, tree-walk statistics, e2e bookkeeping, CPU baseline legs)
against slips that would otherwise only show on the GPU box."""
import argparse
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeSim:
    def __init__(self, c, grav, h):
        import sphb200
        self.S = sphb200
        self.c, self.n, self.grav, self.h = c, len(c["h"]), grav, h
        self.launches = 0

    def sync(self): pass
    def enable_timing(self, on=True): pass
    def launch_count(self): return self.launches
    def stream_ptr(self): return 0
    def fp32_peak_tflops(self): return 72.4
    def close(self): pass
    def upload(self, *a): pass

    def step(self, dt, impl):
        self.launches += 31

    def timings(self):
        g = "gravity_tree" if self.grav == "tree" else "gravity_allpairs"
        return [("smoothing_bounds", 0.04), ("keys_sort_permute_cells", 0.2), ("neighbors_density_eos", 0.8), ("pressure_grad", 0.3),
                (g, 5.0), ("integrate", 0.02)]

    def diagnostics(self):
        return dict(mass=1.0, momentum=np.zeros(3), angular_momentum=np.zeros(3), e_kin=0.0, e_pot=0.0, e_int=0.0,
                    mean_neighbors=52.0, max_neighbors=90)

    def download(self, field, out=None, allow_overflow=False):
        S, n = self.S, self.n
        if field == S.FIELD_SMOOTHING:
            a = np.zeros(n, S.ParticleSmoothing) if out is None else out
            a["influenceArea"] = self.h; a["supportDomain"] = 2 * self.h; a["neighbors"] = 50
            return a
        if field == S.FIELD_GRAVITY:
            if out is not None:
                out[...] = 0
                return out
            a = np.zeros(n, S.GravityField)
            a["numParticles"] = 1100; a["numApprox"] = 420
            return a
        src = {S.FIELD_TRANSLATION: self.c["pos"], S.FIELD_VELOCITY: self.c["vel"], S.FIELD_MASS: self.c["mass"]}.get(field)
        if out is None:
            return src.copy() if src is not None else np.zeros(n, np.float32)
        out[...] = src if src is not None else 0
        return out


class FakeEngine:
    def __init__(self, c, world, rank, local, **params):
        import bench
        self.world, self.rank, self.n, self.cnt, self.b0 = world, rank, len(c["h"]), len(c["h"]), 0
        self.slice = c
        self.sim = FakeSim(c, FakeEngine.grav, np.float32(bench.settled_h_estimate(c)))

    def upload(self, *a): pass
    def step(self, dt, impl): self.sim.step(dt, impl)
    def close(self): pass


class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self, stream=None): pass
    def synchronize(self): pass
    def elapsed_time(self, other): return 12.5


@pytest.mark.parametrize("workload,grav", [("c1", "particle"), ("c2", "tree")])
def test_the_gpu_arm_assembles_a_complete_line(monkeypatch, workload, grav):
    import torch
    import bench
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(bench, "Engine", FakeEngine)
    FakeEngine.grav = grav
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="ours", workload=workload, particles=None, gravity=None, leaf_max=0,
                              aabb_mode=0, kernels_only=False)
    line = bench.measure_workload(args, workload, 1, 0, 0, headline=True)
    d = json.loads(json.dumps(line))                                   # serialisable as it stands
    n = 3000 if workload == "c1" else 10000
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "hbm_passes", "pass_ms", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "sph_particle_steps_per_sec" and d["dtype"] == "f32" and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["ms_per_step"] == 12.5 / 2 and abs(d["value"] - n * 2 / 12.5e-3) < 1e-6
    assert d["gpu_launches"] == 2 * 31
    assert d["config"]["particles"] == n and d["config"]["gravity"] == grav and "l2_policy" in d["config"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == n * 56 and e["d2h_bytes_per_step"] == n * 96 and e["value"] > 0 and e["steps"] == 2
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    pp = d["hbm_passes"]["per_pass"]
    assert set(pp) == {"smoothing_bounds", "keys_sort_permute_cells", "neighbors_density_eos", "pressure_grad", "integrate"}
    assert abs(sum(v["bytes_per_particle"] for k, v in pp.items() if k != "smoothing_bounds") - (368.0 + 12.0 * 52.0)) < 1e-9
    if grav == "tree":
        tw = r["tree_walk"]
        assert tw["interactions_per_particle"] == 1520.0 and tw["ms"] == 5.0
        assert abs(tw["tflops_at_20_flop_per_interaction"] - 20.0 * 1520.0 * n / 5.0e-3 / 1e12) < 1e-9
        assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    else:
        assert r["kernel"] == "k_gravity_allpairs" and r["unit"] == "TFLOP/s"
        assert abs(r["achieved"] - 20.0 * n * (n - 1) / 5.0e-3 / 1e12) < 1e-9 and abs(r["frac"] - r["achieved"] / 72.4) < 1e-12
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["detail"]["extrapolated"] is True
    assert c["detail"]["cell_list_variant"]["particle_steps_per_sec"] > 0
    assert set(c["reference_own_cases"]) == {"c1", "c2"}


def test_a_sub_record_skips_the_cpu_baseline(monkeypatch):
    import torch
    import bench
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(bench, "Engine", FakeEngine)
    FakeEngine.grav = "tree"
    args = argparse.Namespace(gpus=1, steps=1, warmup=0, impl="ours", workload=None, particles=None, gravity=None, leaf_max=0,
                              aabb_mode=0, kernels_only=False)
    d = json.loads(json.dumps(bench.measure_workload(args, "c2", 1, 0, 0, headline=False)))
    assert d["cpu_baseline"]["value"] is None and "skipped" in d["cpu_baseline"]["sample"]
    assert "reference_own_cases" not in d["cpu_baseline"] and d["roofline"]["tree_walk"]["direct_per_particle"] == 1100.0
