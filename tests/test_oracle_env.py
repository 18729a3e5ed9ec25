"""The oracle against what it can be pinned to.

* Philox4x32-10: Random123 known-answer vectors (a real external KAT).
* C oracle vs the committed golden trajectories (tests/golden/traj_*.npz, produced by the
  independent pure-Python restatement oracle/py_env.py with np.logaddexp and scipy's LSA).
* C oracle vs py_env live on fresh seeds.

PARITY UNPINNED vs GS-MARL: the reference's env sources are withheld and it ships no
tests or vectors; these tests pin the oracle to SPEC.md only.
"""
import numpy as np
import pytest

from oracle import gsm_oracle as O, py_env
from tests._util import GOLDEN_TRAJ, golden_path, make_cfg, assert_match, random_actions


def test_philox_known_answers():
    kat = [((0, 0, 0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 6, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for args, want in kat:
        assert tuple(int(x) for x in O.philox(*args)) == want


@pytest.mark.parametrize("name,N,kw", GOLDEN_TRAJ)
def test_c_oracle_matches_golden(name, N, kw):
    z = np.load(golden_path(name, N, kw))
    cfg = make_cfg(name, N, "f64", **kw)
    B, T = z["actions"].shape[1], z["actions"].shape[0]
    env = O.OracleEnv(cfg, B)
    env.set_state(z["agent_state0"], z["landmark_pos"], np.zeros(B, np.int32))
    for t in range(T):
        out = env.step(z["actions"][t])
        want = {k: z[k][t] for k in ("obs", "nbr_idx", "nbr_feat", "nbr_cnt", "adj", "reward",
                                     "cost", "done", "assign")}
        assert_match(out, want, rtol=1e-12, atol=1e-13, ctx=f"{name}{N} t={t}")
        np.testing.assert_allclose(env.agent_state, z["agent_state"][t], rtol=1e-12, atol=1e-13)
    assert z["cost"].sum() > 0          # the fixtures do exercise collisions


@pytest.mark.parametrize("name,N,kw", [("navigation", 5, {}), ("polygon", 7, {"share_reward": True}),
                                       ("line", 3, {}), ("navigation", 2, {"cost_obstacles": False})])
def test_c_oracle_matches_py_env_live(name, N, kw):
    cfg = make_cfg(name, N, "f64", **kw)
    rng = np.random.default_rng(3)
    B = 3
    env = O.OracleEnv(cfg, B)
    env.reset(99)
    env.agent_state *= 0.5
    env.landmark_pos *= 0.5
    pes = [py_env.PyEnv(cfg) for _ in range(B)]
    for b, p in enumerate(pes):
        p.set_state(env.agent_state[b], env.landmark_pos[b])
    for t in range(8):
        a = random_actions(cfg, rng, (B,))
        out = env.step(a)
        for b, p in enumerate(pes):
            po = p.step(a[b])
            assert_match({k: v[b] for k, v in out.items()}, po, rtol=1e-12, atol=1e-13,
                         ctx=f"{name}{N} t={t} b={b}")


def test_reset_is_shard_invariant_and_masked():
    cfg = make_cfg("navigation", 3, "f64")
    full = O.OracleEnv(cfg, 8)
    full.reset(42)
    lo = O.OracleEnv(cfg, 3, env_offset=0)
    hi = O.OracleEnv(cfg, 5, env_offset=3)
    lo.reset(42)
    hi.reset(42)
    assert (np.concatenate([lo.agent_state, hi.agent_state]) == full.agent_state).all()
    assert (np.concatenate([lo.landmark_pos, hi.landmark_pos]) == full.landmark_pos).all()
    before = full.agent_state.copy()
    mask = np.array([1, 0, 0, 1, 0, 0, 0, 0], np.uint8)
    full.reset(42, mask)
    assert (full.agent_state[1:3] == before[1:3]).all() and (full.agent_state[4:] == before[4:]).all()
    assert (full.agent_state[0] != before[0]).any()        # new episode -> new draw
    assert (full.episode == np.array([2, 1, 1, 2, 1, 1, 1, 1])).all()
    ext = cfg.spawn_extent[0]
    assert (np.abs(full.agent_state[..., :2]) <= ext).all() and (full.agent_state[..., 2:] == 0).all()


def test_f32_oracle_tracks_f64_one_step():
    cfg64 = make_cfg("navigation", 6, "f64")
    cfg32 = cfg64.replace(dtype="f32")
    e64, e32 = O.OracleEnv(cfg64, 16), O.OracleEnv(cfg32, 16)
    e64.reset(5)
    e32.set_state(e64.agent_state, e64.landmark_pos)
    a = random_actions(cfg64, np.random.default_rng(0), (16,))
    o64, o32 = e64.step(a), e32.step(a)
    np.testing.assert_allclose(o32["obs"], o64["obs"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(o32["reward"], o64["reward"], rtol=1e-4, atol=1e-5)


def test_threads_give_identical_results():
    cfg = make_cfg("navigation", 4, "f64")
    a = random_actions(cfg, np.random.default_rng(1), (64,))
    outs = []
    for nt in (1, 4):
        O.set_threads(nt)
        e = O.OracleEnv(cfg, 64)
        e.reset(8)
        outs.append((e.step(a), e.agent_state.copy()))
    O.set_threads(1)
    for k in outs[0][0]:
        assert (outs[0][0][k] == outs[1][0][k]).all()
    assert (outs[0][1] == outs[1][1]).all()


def test_closed_form_ballistic_motion_and_clamp():
    """SPEC §2/§4 without contacts: v' = v(1-damping) + accel*u/mass*dt (then clamped), p' = p + v'dt."""
    cfg = make_cfg("navigation", 3, "f64", n_obstacles=0, max_speed=[0.0, 0.5, 0.0])
    e = O.OracleEnv(cfg, 1)
    ag = np.array([[[-3.0, 0.0, 0.2, -0.1], [0.0, 3.0, 0.0, 0.0], [3.0, -3.0, -0.4, 0.3]]])
    lm = np.array([[[5.0, 5.0], [6.0, 6.0], [7.0, 7.0]]])
    e.set_state(ag, lm, [0])
    u = np.array(cfg.discrete_u)
    acts = np.array([[1, 3, 4]])
    e.step(acts)
    for i in range(3):
        v = ag[0, i, 2:] * (1 - cfg.damping) + (cfg.accel[i] * u[acts[0, i]] / cfg.mass[i]) * cfg.dt
        ms = cfg.max_speed[i]
        if ms > 0 and np.hypot(*v) > ms:
            v = v / np.hypot(*v) * ms
        np.testing.assert_allclose(e.agent_state[0, i, 2:], v, rtol=1e-15, atol=1e-18)
        np.testing.assert_allclose(e.agent_state[0, i, :2], ag[0, i, :2] + v * cfg.dt, rtol=1e-15, atol=1e-18)
    assert np.isclose(np.hypot(*e.agent_state[0, 1, 2:]), 0.5)       # agent 1 hit its speed clamp
    assert e.step_count[0] == 1


def test_contact_forces_are_equal_and_opposite_and_match_softplus():
    """SPEC §3: two overlapping agents, no action: the impulses are exactly opposite and equal to
    contact_force * softplus(-(d - dmin)/margin) * margin along the line of centres."""
    cfg = make_cfg("navigation", 2, "f64", n_obstacles=0)
    e = O.OracleEnv(cfg, 1)
    d = 0.15                                                          # dmin = 0.20 -> 0.05 overlap
    ag = np.array([[[0.0, 0.0, 0.0, 0.0], [d * 0.6, d * 0.8, 0.0, 0.0]]])
    lm = np.array([[[9.0, 9.0], [9.0, -9.0]]])
    e.set_state(ag, lm, [0])
    out = e.step(np.array([[0, 0]]))
    v0, v1 = e.agent_state[0, 0, 2:], e.agent_state[0, 1, 2:]
    assert (v0 == -v1).all()                                         # bit-exact action = reaction
    k = cfg.contact_margin
    pen = np.logaddexp(0.0, -(d - 0.2) / k) * k
    f = cfg.contact_force * pen                                       # magnitude
    np.testing.assert_allclose(np.hypot(*v1), f / cfg.mass[1] * cfg.dt, rtol=1e-12)
    np.testing.assert_allclose(v1 / np.hypot(*v1), [0.6, 0.8], rtol=1e-12)
    # the pair separated by exactly cf*dt^2/m * pen along the axis; cost counted before? no: after
    assert out["cost"].tolist() == [[1.0, 1.0]] or out["cost"].tolist() == [[0.0, 0.0]]
    new_d = np.hypot(*(e.agent_state[0, 1, :2] - e.agent_state[0, 0, :2]))
    assert (new_d < 0.2) == bool(out["cost"][0, 0])                  # cost is taken on the post-step state


def test_oracle_reproduces_goldens_without_the_product_package(tmp_path):
    """the oracle builds its worlds from its OWN literal tables (oracle/worlds.py).
    With `gs_marl_b200` made un-importable, tests/golden/make_golden.py regenerates every committed
    trajectory bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    blocker = (
        "import sys, importlib.abc\n"
        "class B(importlib.abc.MetaPathFinder):\n"
        "    def find_spec(self, name, path=None, target=None):\n"
        "        if name == 'gs_marl_b200' or name.startswith('gs_marl_b200.'):\n"
        "            raise ImportError('product package blocked for this test')\n"
        "sys.meta_path.insert(0, B())\n"
        f"sys.argv = ['make_golden.py', '--traj-only', {str(tmp_path)!r}]\n"
        "import runpy\n"
        f"runpy.run_path({os.path.join(root, 'tests', 'golden', 'make_golden.py')!r}, run_name='__main__')\n"
        "assert not any(m.startswith('gs_marl_b200') for m in sys.modules)\n")
    subprocess.run([sys.executable, "-c", blocker], check=True, cwd=root, timeout=600)
    for name, N, kw in GOLDEN_TRAJ:
        want, got = np.load(golden_path(name, N, kw)), np.load(os.path.join(tmp_path, os.path.basename(golden_path(name, N, kw))))
        assert set(want.files) == set(got.files)
        for k in want.files:
            assert np.array_equal(want[k], got[k]), (name, N, k)


def test_oracle_tables_match_product_scenarios():
    """The two independently written world tables (oracle/worlds.py, gs_marl_b200/scenarios + presets)
    agree for every BASELINE.json configuration and every golden trajectory."""
    from tests._util import assert_same_world, oracle_world
    from gs_marl_b200 import scenarios
    cases = list(GOLDEN_TRAJ) + [("navigation", n, {}) for n in (3, 6, 12)] + \
        [("navigation", n, {"max_nbrs": 32}) for n in (24, 48, 96)] + \
        [(s, n, {}) for s in ("polygon", "line", "simple_formation", "simple_line") for n in (3, 6, 12)]
    for name, N, kw in cases:
        for dt in ("f32", "f64"):
            assert_same_world(scenarios.load(name).make_world(N, dtype=dt, **kw), oracle_world(name, N, dt, **kw),
                              f"{name}-{N}")
