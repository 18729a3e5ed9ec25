"""tools/unblock.py — the kit that turns "the reference sources appeared" into goldens, presets and a
per-section SPEC.md diff in one command — proven end to end on the synthetic
stand-in tree of tools/fake_gsmarl.py, plus the replay of its recordings through the C oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import gsm_oracle as O
from tests._util import load_ref_golden, ref_golden_files, ref_world_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fake_gsmarl  # noqa: E402


def _kit(ref, out, *extra):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "unblock.py"), "--ref", str(ref), "--out", str(out),
                           *extra], capture_output=True, text=True, timeout=900, cwd=ROOT)


def test_gate_reports_blocked_when_the_hot_path_is_absent(tmp_path):
    """A tree like today's /root/reference: egg-info manifest, no gsmarl/ -> BLOCKED, exit code 2."""
    ref = tmp_path / "ref"
    (ref / "GSMARL.egg-info").mkdir(parents=True)
    (ref / "GSMARL.egg-info" / "SOURCES.txt").write_text(fake_gsmarl.FILES["GSMARL.egg-info/SOURCES.txt"])
    p = _kit(ref, tmp_path / "out")
    assert p.returncode == 2 and "BLOCKED" in p.stdout
    g = json.load(open(tmp_path / "out" / "gate.json"))
    assert g["blocked"] and g["present"] == 0 and "gsmarl/envs/mpe_env/multiagent/core.py" in g["hot_path_missing"]


def test_kit_end_to_end_on_the_synthetic_tree(tmp_path):
    fake_gsmarl.write_tree(str(tmp_path / "ref"))
    p = _kit(tmp_path / "ref", tmp_path / "out")
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    rep = json.load(open(tmp_path / "out" / "report.json"))
    assert not rep["blocked"] and any("gym" in s for s in rep["shims"])      # imported under the gym.spaces shim
    assert set(rep["configs"]) == {"nav-3", "nav-6", "nav-12", "nav-24", "polygon-6", "polygon-12", "line-6", "line-12"}
    for label, c in rep["configs"].items():
        assert c["status"] == "ok", (label, c)
        sections = {r["section"] for r in c["rows"]}
        assert {"§2-4", "§6", "§7"} <= sections and (label.startswith("nav") or "§5" in sections)
        assert all(r["status"] == "PASS" for r in c["rows"]), (label, [r for r in c["rows"] if r["status"] != "PASS"])
        assert c["unmapped_constants"] == [] and c["changed_constants"] == [], (label, c)
        assert os.path.exists(tmp_path / "out" / "goldens" / f"ref_{label}.npz")
        preset = json.load(open(tmp_path / "out" / "presets" / f"{label}.json"))
        assert preset["world"]["damping"] == 0.25 and len(preset["agents"]) == int(label.split("-")[1])
    assert "| nav-3 | §2-4 |" in open(tmp_path / "out" / "SPEC_DIFF.md").read()


def test_kit_flags_the_section_that_differs(tmp_path):
    """One changed constant, one changed physics rule, one changed reward rule in the tree -> the diff
    table says CHANGED for the constant, FAIL for SPEC §2-4 positions and §7 reward, PASS elsewhere."""
    fake_gsmarl.write_tree(str(tmp_path / "ref"), damping=0.5, reward_bug=True, physics_bug=True)
    p = _kit(tmp_path / "ref", tmp_path / "out", "--configs", "nav-3,polygon-6")
    assert p.returncode == 1                                  # FAIL rows -> non-zero exit
    rep = json.load(open(tmp_path / "out" / "report.json"))
    for label in ("nav-3", "polygon-6"):
        c = rep["configs"][label]
        st = {r["what"].split(" (")[0].split("(")[0].strip(): r["status"] for r in c["rows"]}
        assert st["positions after step"] == "FAIL" and st["velocities after step"] == "PASS"
        assert st["reward"] == "FAIL" and st["cost"] == "PASS" and st["observation"] == "PASS"
        assert st["neighbour sets"] == "PASS" and st["padded neighbour rows"] == "PASS"
        assert [x["field"] for x in c["changed_constants"]] == ["damping"]
        assert c["changed_constants"][0]["reference"] == "0.5"


@pytest.fixture(scope="module")
def ref_files(tmp_path_factory):
    return ref_golden_files(tmp_path_factory)


def test_ref_golden_replay_c_oracle(ref_files):
    """Every recorded reference transition replayed through the C oracle (fp64): post-step state
    within 1e-9, observation / reward within 1e-9, cost exact."""
    assert ref_files
    for path in ref_files:
        rec, fields = load_ref_golden(path)
        _, world = ref_world_pair(fields)
        T = rec["state_before"].shape[0]
        env = O.OracleEnv(world, T)
        env.set_state(rec["state_before"], rec["landmarks"], np.zeros(T, np.int32))
        out = env.step(rec["control"])
        np.testing.assert_allclose(env.agent_state, rec["state_after"], rtol=0, atol=1e-9, err_msg=path)
        np.testing.assert_allclose(out["obs"], rec["obs_cb"], rtol=0, atol=1e-9, err_msg=path)
        np.testing.assert_allclose(out["reward"], rec["reward_cb"][..., 0], rtol=0, atol=1e-9, err_msg=path)
        assert (out["cost"] == rec["cost_cb"][..., 0]).all(), path
