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
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-budget", "2"], 300)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert d["unit"] == "agent-steps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "withheld" in d["cpu_baseline"]["sample"]   # never presented as GS-MARL's own env
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0 and "unpinned" in d["config"]["spec_status"]


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "200", "--warmup", "25", "--e2e-steps", "10", "--closed-loop-steps", "25",
              "--large-envs", "65536", "--cpu-budget", "2"], 600)
    assert BASE_KEYS | {"clocks", "roofline"} <= set(d)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.05 < r["frac"] < 1.2 and r["traffic"] is None or r["traffic"] > 0
    assert d["gpu_launches"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 13221888
    assert d["e2e"]["value"] < d["value"]              # the host path cannot beat the device path
    assert d["cpu_baseline"]["kind"] == "port" and d["dtype"] == "f32" and d["scaling"] == "weak"
    assert d["config"]["workload"].startswith("cooperative navigation, 3 agents, 16384 envs")
    e = d["e2e"]
    assert 0 < e["d2h_gbs_achieved"] < e["d2h_gbs_ceiling"] * 1.05      # below the measured pinned-copy rate
    f = d["closed_loop"]["fused_actor"]
    assert f["value"] > d["closed_loop"]["value"] and f["actor_kernel_us"] > 0
