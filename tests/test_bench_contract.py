"""bench.py's output contract: the CPU reference arm runs here end to end; the GPU arm's line
is checked on the B200 box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True,
                         text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                      # exactly one JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--envs", "256"], 300)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert d["unit"] == "agent-steps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "withheld" in d["cpu_baseline"]["sample"]   # never presented as GS-MARL's own env
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0 and "unpinned" in d["config"]["spec_status"]
    # the reference arm never shrinks its workload: exactly the envs it was asked for, and the same
    # `config` object the GPU arm prints for the same command line
    assert d["config"]["envs_per_gpu"] == 256 and d["cpu_baseline"]["sample_envs_per_step"] == 256
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(1, 256)


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "20", "--warmup", "5", "--e2e-steps", "10", "--closed-loop-steps", "25",
              "--large-envs", "65536", "--cpu-budget", "2", "--region-ms", "100", "--config-ms", "20",
              "--config5-steps", "25"], 900)
    assert BASE_KEYS | {"clocks", "roofline", "repeats", "configs", "method"} <= set(d)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(1)       # same object as the reference arm's
    # the timed region: `repeats` K-step regions replayed from a CUDA graph, >= region-ms long, and
    # value / ms_per_step / roofline all derived from that one region
    assert d["steps"] == 20 and d["repeats"] >= 1 and d["timed_region_ms"] >= 100
    assert abs(d["ms_per_step"] - d["timed_region_ms"] / (d["repeats"] * d["steps"])) < 1e-9
    assert abs(d["value"] - 16384 * 3 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    r = d["roofline"]
    assert abs(r["step_us"] - d["ms_per_step"] * 1e3) < 1e-6
    assert abs(r["achieved"] - r["algorithmic_bytes_per_agent_step"] * d["value"] / 1e9) / r["achieved"] < 1e-6
    assert r["frac"] < r["frac_per_step_accounting"]      # state / landmarks / counter charged once per launch
    # (a cold nvidia-smi can deliver its first sample after a 100 ms region: the sampler then reports the
    # warm-up + region window and says so)
    assert d["clocks"]["window"].endswith("timed region") or "timed region" in d["clocks"]["window"]
    assert d["clocks"]["samples"] >= 1
    labels = [c["config"] for c in d["configs"]]
    assert sum(l.startswith("configs[2]") for l in labels) == 3 and sum(l.startswith("configs[3]") for l in labels) == 4
    assert sum(l.startswith("configs[4]") for l in labels) == 1
    for c in d["configs"]:
        assert c["step_us"] > 0 and c["agent_steps_per_s"] > 0
        if not c["config"].startswith("configs[4]"):
            assert 0.02 < c["frac"] < 1.2 and c["bytes_per_agent_step"] > 0
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.05 < r["frac"] < 1.2 and r["traffic"] is None or r["traffic"] > 0
    assert d["gpu_launches"] > 0 and d["e2e"]["d2h_bytes_per_step_dense"] == 13221888 and 0 < d["e2e"]["d2h_bytes_per_step"] < 13221888
    assert d["e2e"]["value"] < d["value"]              # the host path cannot beat the device path
    assert d["cpu_baseline"]["kind"] == "port" and d["dtype"] == "f32" and d["scaling"] == "weak"
    assert d["config"]["workload"].startswith("cooperative navigation, 3 agents, 16384 envs")
    e = d["e2e"]
    assert 0 < e["d2h_gbs_achieved"] < e["d2h_gbs_ceiling"] * 1.25      # about or below the measured pinned-copy rate (10 steps: noisy)
    f = d["closed_loop"]["fused_actor"]
    assert f["value"] > d["closed_loop"]["value"] and f["actor_kernel_us"] > 0
