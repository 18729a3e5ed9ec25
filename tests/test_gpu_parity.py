"""Parity of the CUDA path (through the C ABI) with the oracle — the `-m gpu` gate.

Bar (BASELINE.json north_star): fp64 verification mode — neighbour sets, adjacency masks,
collision/cost counts, done flags and assignments BIT-EXACT, real outputs within 1e-9 over
25 steps; fp32 production mode — within 1e-4 relative.  "Oracle" = SPEC.md restated on the
CPU (parity with GS-MARL itself is unpinned: its env sources are withheld).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from gs_marl_b200 import abi
from tests._util import (GOLDEN_TRAJ, golden_path, make_cfg, assert_match, random_actions, GOLDEN,
                         near_threshold_rows)

pytestmark = pytest.mark.gpu

F64_RTOL, F64_ATOL = 1e-9, 1e-9
OUT_KEYS = ("obs", "nbr_idx", "nbr_feat", "nbr_cnt", "adj", "reward", "cost", "done", "assign")


def _np(bufs):
    out = {}
    for k, v in bufs.items():
        a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        out[k] = a.view(np.uint32) if k == "adj" else a
    return out


def _env(cfg, n, **kw):
    from gs_marl_b200.environment import MultiAgentGraphConstrainEnv
    return MultiAgentGraphConstrainEnv(cfg, n, **kw)


def _squeezed_start(cfg, n, seed, scale=0.45):
    from oracle import gsm_oracle as O
    o = O.OracleEnv(cfg, n)
    o.reset(seed)
    o.agent_state *= scale
    o.landmark_pos *= scale
    return o


@pytest.fixture(params=[None, 1, 2, 4, 8, "cta32", "cta8"])
def mapping(request, monkeypatch):
    """Every thread mapping of the env kernel must give the same answers."""
    m = request.param
    if m is None:
        return m
    if isinstance(m, int):
        monkeypatch.setenv("GSM_FORCE_CTA_ENV", "0")
        monkeypatch.setenv("GSM_FORCE_P", str(m))
    else:
        monkeypatch.setenv("GSM_FORCE_CTA_ENV", "1")
        monkeypatch.setenv("GSM_FORCE_P", m[3:])
    return m


# ---- golden fixtures ------------------------------------------------------------------
@pytest.mark.parametrize("name,N,kw", GOLDEN_TRAJ)
def test_f64_matches_golden_trajectories(name, N, kw, mapping):
    if isinstance(mapping, int) and N * mapping > 32:
        pytest.skip("packed mapping needs N*P <= 32")
    z = np.load(golden_path(name, N, kw))
    cfg = make_cfg(name, N, "f64", **kw)
    T, B = z["actions"].shape[:2]
    env = _env(cfg, B)
    env.set_state(z["agent_state0"], z["landmark_pos"], np.zeros(B, np.int32))
    for t in range(T):
        env.step(z["actions"][t])
        got = _np(env.buf)
        want = {k: z[k][t] for k in OUT_KEYS}
        assert_match(got, want, rtol=F64_RTOL, atol=F64_ATOL, ctx=f"{name}{N} t={t} map={mapping}")
        ag = env.get_state()[0].cpu().numpy()
        np.testing.assert_allclose(ag, z["agent_state"][t], rtol=F64_RTOL, atol=F64_ATOL)
    env.close()


# ---- oracle on seeded inputs, every scenario / size the configs name --------------------
CASES = [("navigation", 3, 512, {}), ("navigation", 6, 128, {}), ("navigation", 12, 64, {}),
         ("navigation", 24, 16, {"max_nbrs": 16}), ("navigation", 48, 6, {"max_nbrs": 32}),
         ("navigation", 96, 3, {"max_nbrs": 32}),
         ("polygon", 3, 128, {}), ("polygon", 6, 128, {}), ("polygon", 12, 64, {"share_reward": True}),
         ("line", 6, 128, {}), ("line", 12, 64, {}),
         ("navigation", 5, 40, {"action_mode": "continuous", "share_reward": True}),
         ("navigation", 1, 33, {"n_obstacles": 0, "max_nbrs": 1}),
         ("navigation", 32, 5, {"n_obstacles": 3, "max_nbrs": 8}),
         ("navigation", 33, 5, {"n_obstacles": 0, "max_nbrs": 65})]


@pytest.mark.parametrize("name,N,B,kw", CASES)
def test_f64_25_steps_vs_oracle(name, N, B, kw):
    cfg = make_cfg(name, N, "f64", **kw)
    o = _squeezed_start(cfg, B, 1234 + N)
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    rng = np.random.default_rng(N)
    total_cost = 0.0
    for t in range(25):
        a = random_actions(cfg, rng, (B,))
        want = o.step(a)
        env.step(a)
        assert_match(_np(env.buf), want, rtol=F64_RTOL, atol=F64_ATOL, ctx=f"{name}{N} t={t}")
        total_cost += want["cost"].sum()
    ag, lm, tt = env.get_state()
    np.testing.assert_allclose(ag.cpu().numpy(), o.agent_state, rtol=F64_RTOL, atol=F64_ATOL)
    assert (tt.cpu().numpy() == o.step_count).all()
    assert _np(env.buf)["done"].all() == (25 >= cfg.episode_length)
    if N >= 3:
        assert total_cost > 0, "case never exercised a collision"
    env.close()


@pytest.mark.parametrize("name,N,P", [("navigation", 3, 8), ("navigation", 3, 4), ("navigation", 3, 2),
                                      ("navigation", 3, 1), ("navigation", 6, 4), ("navigation", 6, 1),
                                      ("polygon", 6, 4), ("polygon", 6, 1), ("line", 6, 4), ("line", 6, 1),
                                      ("polygon", 3, 1), ("polygon", 5, 1), ("line", 4, 1), ("polygon", 12, 1),
                                      ("polygon", 12, 2), ("line", 12, 1), ("line", 12, 2), ("navigation", 12, 2)])
@pytest.mark.parametrize("kw", [{}, {"max_nbrs": 3, "share_reward": True}])
def test_specialised_kernel_variants_f64(name, N, P, kw, monkeypatch):
    """Every compiled (scenario, N, L, P) instance of the register-resident kernel, single
    steps and one fused 25-step launch, against the oracle."""
    monkeypatch.setenv("GSM_SPEC_P", str(P))
    cfg = make_cfg(name, N, "f64", **kw)
    B, T = 77, 25
    o = _squeezed_start(cfg, B, 31 + N)
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    rng = np.random.default_rng(P)
    acts = random_actions(cfg, rng, (T, B))
    wants = [{k: v.copy() for k, v in o.step(acts[t]).items()} for t in range(T)]
    roll = _np(env.rollout(acts))
    for t in range(T):
        assert_match({k: roll[k][t] for k in OUT_KEYS}, wants[t], rtol=F64_RTOL, atol=F64_ATOL,
                     ctx=f"{name}{N} P={P} fused t={t}")
    env.set_state(o.agent_state * 0 + _squeezed_start(cfg, B, 31 + N).agent_state, o.landmark_pos, np.zeros(B, np.int32))
    for t in range(3):
        env.step(acts[t])
        assert_match(_np(env.buf), wants[t], rtol=F64_RTOL, atol=F64_ATOL, ctx=f"{name}{N} P={P} step t={t}")
    env.close()


@pytest.mark.parametrize("N,K,B,kw", [(24, 16, 9, {}), (48, 32, 5, {"share_reward": True}),
                                      (96, 32, 3, {"action_mode": "continuous"}), (13, 4, 11, {}),
                                      (12, 32, 37, {}), (24, 7, 6, {"own_goal_always": False}),
                                      (33, 64, 4, {"n_obstacles": 0, "cost_obstacles": False})])
def test_large_team_kernel_fused_f64(N, K, B, kw, monkeypatch):
    """env_lane_kernel (lane per agent; teams of 12 ... 96, one or several warps per env): one fused
    10-step launch and single steps against the oracle."""
    monkeypatch.setenv("GSM_LANE_MIN_N", "12")
    cfg = make_cfg("navigation", N, "f64", max_nbrs=K, **kw)
    T = 10
    o = _squeezed_start(cfg, B, 5 + N)
    s0 = o.agent_state.copy()
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    rng = np.random.default_rng(N)
    acts = random_actions(cfg, rng, (T, B))
    wants = [{k: v.copy() for k, v in o.step(acts[t]).items()} for t in range(T)]
    roll = _np(env.rollout(acts))
    assert env.kernel_launches == 1                      # fused, not a graph of T launches
    for t in range(T):
        assert_match({k: roll[k][t] for k in OUT_KEYS}, wants[t], rtol=F64_RTOL, atol=F64_ATOL,
                     ctx=f"nav{N} lane fused t={t}")
    np.testing.assert_allclose(env.get_state()[0].cpu().numpy(), o.agent_state, rtol=F64_RTOL, atol=F64_ATOL)
    env.set_state(s0, o.landmark_pos, np.zeros(B, np.int32))
    for t in range(2):
        env.step(acts[t])
        assert_match(_np(env.buf), wants[t], rtol=F64_RTOL, atol=F64_ATOL, ctx=f"nav{N} lane step t={t}")
    env.close()


@pytest.mark.parametrize("name,N,B,kw", CASES)
def test_f32_one_step_within_1e4(name, N, B, kw):
    """Production precision: one step from identical (fp32-representable) states."""
    cfg64 = make_cfg(name, N, "f64", **kw)
    cfg32 = cfg64.replace(dtype="f32")
    o = _squeezed_start(cfg64, B, 77 + N)
    o.agent_state[...] = o.agent_state.astype(np.float32)
    o.landmark_pos[...] = o.landmark_pos.astype(np.float32)
    env = _env(cfg32, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    a = random_actions(cfg64, np.random.default_rng(N), (B,))
    want = o.step(a)
    env.step(a)
    got = _np(env.buf)
    ag = env.get_state()[0].cpu().numpy()
    np.testing.assert_allclose(ag, o.agent_state, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(got["obs"], want["obs"], rtol=1e-4, atol=1e-5)
    # integer outputs: bit-identical wherever no fp64 predicate sits within 1e-5 of its threshold
    # (the UNVERIFIED preset has contact_force*dt^2/mass == 1, so resting contacts land exactly
    # on dist == dmin and fp32 may round the collision flag either way)
    near = near_threshold_rows(cfg64, o.agent_state, o.landmark_pos, 1e-5)
    assert near.mean() < 0.05
    ok = ~near
    for k in ("nbr_cnt", "cost", "nbr_idx", "adj"):
        g, w = got[k][ok], want[k][ok]
        assert (g == w).all(), (k, int((g != w).sum()))
    ok = ok & (got["assign"] == want["assign"])
    np.testing.assert_allclose(got["reward"][ok], want["reward"][ok], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(got["nbr_feat"][ok], want["nbr_feat"][ok], rtol=1e-4, atol=1e-5)
    env.close()


def test_f32_matches_f32_oracle_25_steps():
    """fp32 CUDA vs the oracle run in fp32 (same rounding except libm/FMA): trajectories stay
    within 1e-4 relative over 25 steps."""
    from oracle import gsm_oracle as O
    cfg = make_cfg("navigation", 3, "f32")
    B = 256
    o = O.OracleEnv(cfg, B)
    o.reset(5)
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    rng = np.random.default_rng(0)
    for t in range(25):
        a = random_actions(cfg, rng, (B,))
        want = o.step(a)
        env.step(a)
    got = _np(env.buf)
    np.testing.assert_allclose(got["obs"], want["obs"], rtol=1e-4, atol=2e-4)
    assert (got["nbr_cnt"] == want["nbr_cnt"]).mean() > 0.99
    env.close()


def _contact_touched(cfg, agent_state, landmark_pos, margin):
    """Bool [n_envs, N]: the agent has a colliding pair within `margin` of contact (the stiff contact
    force amplifies rounding differences of that agent's state from here on), and the agent-agent
    distances [n_envs, N, N]."""
    N = cfg.n_agents
    pos = np.concatenate([agent_state[..., :2], landmark_pos], axis=1).astype(np.float64)
    d = np.sqrt(((pos[:, None, :, :] - pos[:, :N, None, :]) ** 2).sum(-1))                    # [B, N, E]
    size, col = np.asarray(cfg.size, np.float64), np.asarray(cfg.collide, bool)
    hit = (d < (size[:N, None] + size[None, :])[None] + margin) & (col[:N, None] & col[None, :])[None]
    for i in range(N):
        hit[:, i, i] = False
    return hit.any(2), d[:, :, :N]


@pytest.mark.parametrize("name,N,B,kw", [("navigation", 3, 512, {}), ("navigation", 12, 128, {}),
                                          ("navigation", 24, 64, {"max_nbrs": 32}), ("navigation", 96, 8, {"max_nbrs": 32}),
                                          ("polygon", 6, 256, {}), ("polygon", 12, 128, {}), ("line", 12, 128, {})])
def test_f32_fused_rollout_25_steps_vs_f32_oracle(name, N, B, kw):
    """Production precision over a whole FUSED 25-step launch (specialised, lane and team kernels)
    against the oracle run in fp32, ALL outputs, every step.  Contacts are stiff (SPEC §3:
    contact_force / contact_margin), so an AGENT is compared up to the step where one of its
    colliding pairs first comes within 0.03 of touching (its own state is exact to rounding until
    then: the contact term is cut / < e^-30 beyond that margin); its row is compared while every
    agent within sensing radius + 0.5 is such a clean agent, so every neighbour row and every
    neighbour predicate rests on accurate positions.  Rows with a predicate within 1e-4 of its
    threshold are excluded from the integer comparison (SPEC §9 allows those to flip).  Outputs
    that depend on the whole env (shared reward, the assignment and what follows from it) are
    compared in envs whose agents are all clean."""
    from oracle import gsm_oracle as O
    cfg = make_cfg(name, N, "f32", **kw)
    T = 25
    lsa = name != "navigation"
    o = O.OracleEnv(cfg, B)
    o.reset(11 + N)
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    rng = np.random.default_rng(N)
    acts = random_actions(cfg, rng, (T, B))
    out = _np(env.rollout(acts))
    touched, _ = _contact_touched(cfg, o.agent_state, o.landmark_pos, 0.03)
    clean = ~touched                                                                          # [B, N], cumulative
    rows_checked = assign_bad = env_rows = 0
    for t in range(T):
        want = o.step(acts[t])
        touched, d_aa = _contact_touched(cfg, o.agent_state, o.landmark_pos, 0.03)
        clean &= ~touched
        dirty_near = ((d_aa < cfg.sensing_radius + 0.5) & ~clean[:, None, :]).any(2)
        near = near_threshold_rows(cfg, o.agent_state, o.landmark_pos, 1e-4)
        env_clean = clean.all(1)
        ok = clean & ~dirty_near & ~near
        for k in ("nbr_cnt", "cost", "nbr_idx", "adj", "done"):
            g, w = out[k][t][ok], want[k][ok]
            assert (g == w).all(), (name, N, t, k, int((g != w).sum()))
        np.testing.assert_allclose(out["nbr_feat"][t][ok], want["nbr_feat"][ok], rtol=1e-4, atol=2e-5,
                                   err_msg=f"{name}{N} t={t} nbr_feat")
        np.testing.assert_allclose(out["obs"][t][ok][:, :4], want["obs"][ok][:, :4], rtol=1e-4, atol=2e-5,
                                   err_msg=f"{name}{N} t={t} obs[:4]")
        # the target-relative part of obs and the reward follow the assignment (polygon / line)
        tgt_ok = ok & env_clean[:, None] if lsa else ok
        same_asg = out["assign"][t] == want["assign"]
        assign_bad += int((tgt_ok & ~same_asg).sum())
        env_rows += int(tgt_ok.sum())
        tgt_ok = tgt_ok & same_asg
        np.testing.assert_allclose(out["obs"][t][tgt_ok], want["obs"][tgt_ok], rtol=1e-4, atol=2e-5,
                                   err_msg=f"{name}{N} t={t} obs")
        rew_ok = tgt_ok & env_clean[:, None] if cfg.share_reward else tgt_ok
        np.testing.assert_allclose(out["reward"][t][rew_ok], want["reward"][rew_ok], rtol=1e-4, atol=2e-5,
                                   err_msg=f"{name}{N} t={t} reward")
        rows_checked += int(ok.sum())
    assert rows_checked > 0.05 * T * B * N, "too few rows survived the contact / threshold masks"
    if lsa:
        assert env_rows > 0.02 * T * B * N, "too few whole-env rows for the assignment comparison"
    assert assign_bad <= 0.01 * max(env_rows, 1)
    env.close()


def test_ref_golden_replay_cuda(tmp_path_factory):
    """The CUDA kernels (fp64) on trajectories RECORDED FROM A REFERENCE TREE by tools/unblock.py:
    GSM_REF_GOLDEN_DIR once the real gsmarl/ is mounted; until then recordings of the synthetic
    stand-in tree (tools/fake_gsmarl.py), an implementation that shares no code with the product or
    the oracle.  Every recorded transition is one env: state in, control in, state / obs / reward
    within 1e-9 and cost exact."""
    from tests._util import load_ref_golden, ref_golden_files, ref_world_pair
    files = ref_golden_files(tmp_path_factory)
    assert files
    for path in files:
        rec, fields = load_ref_golden(path)
        cfg, _ = ref_world_pair(fields)
        T = rec["state_before"].shape[0]
        env = _env(cfg, T)
        env.set_state(rec["state_before"], rec["landmarks"], np.zeros(T, np.int32))
        env.step(rec["control"])
        got = _np(env.buf)
        ag = env.get_state()[0].cpu().numpy()
        np.testing.assert_allclose(ag, rec["state_after"], rtol=0, atol=1e-9, err_msg=path)
        np.testing.assert_allclose(got["obs"], rec["obs_cb"], rtol=0, atol=1e-9, err_msg=path)
        np.testing.assert_allclose(got["reward"], rec["reward_cb"][..., 0], rtol=0, atol=1e-9, err_msg=path)
        assert (got["cost"] == rec["cost_cb"][..., 0]).all(), path
        env.close()


# ---- reset (SPEC §8): integer RNG -> bit-exact states -----------------------------------
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_reset_bit_exact_and_shard_invariant(dtype):
    from oracle import gsm_oracle as O
    cfg = make_cfg("navigation", 3, dtype)
    o = O.OracleEnv(cfg, 300)
    o.reset(2026)
    env = _env(cfg, 300, seed=2026)
    obs, graph = env.reset()
    ag, lm, t = env.get_state()
    assert (ag.cpu().numpy() == o.agent_state).all() and (lm.cpu().numpy() == o.landmark_pos).all()
    want = o.observe()
    tol = dict(rtol=F64_RTOL, atol=F64_ATOL) if dtype == "f64" else dict(rtol=1e-6, atol=1e-6)
    assert_match(_np(env.buf), want, ctx="reset obs", keys=("obs", "nbr_idx", "nbr_feat", "nbr_cnt", "adj", "assign"), **tol)
    # a shard starting at env 100 draws the same worlds
    sh = _env(cfg, 50, env_offset=100, seed=2026)
    sh.reset()
    assert (sh.get_state()[0].cpu().numpy() == o.agent_state[100:150]).all()
    # masked re-reset: second episode for the masked envs only
    mask = np.zeros(300, np.uint8)
    mask[::7] = 1
    o.reset(2026, mask)
    env.reset(torch.as_tensor(mask))
    ag2 = env.get_state()[0].cpu().numpy()
    assert (ag2 == o.agent_state).all()
    assert (ag2[1] == ag.cpu().numpy()[1]).all() and (ag2[0] != ag.cpu().numpy()[0]).any()
    env.close(); sh.close()


# ---- stand-alone batched LSA kernel ---------------------------------------------------------
def _gsm_lsa(cost):
    lib = abi.load_library()
    c = torch.as_tensor(cost).cuda().contiguous()
    B, n = c.shape[0], c.shape[-1]
    out = torch.full((B, n), -7, dtype=torch.int32, device="cuda")
    st = lib.gsm_lsa(c.data_ptr(), out.data_ptr(), B, n, abi.GSM_F64 if c.dtype == torch.float64 else abi.GSM_F32,
                     0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == 0, lib.gsm_last_error(None)
    return out.cpu().numpy()


def test_lsa_kernel_matches_scipy_golden():
    z = np.load(os.path.join(GOLDEN, "lsa_scipy.npz"))
    by_n = {}
    for k in range(int(z["count"])):
        by_n.setdefault(z[f"cost_{k}"].shape[0], []).append(k)
    for n, ks in by_n.items():
        cost = np.stack([z[f"cost_{k}"] for k in ks])
        want = np.stack([z[f"col4row_{k}"] for k in ks])
        got = _gsm_lsa(cost)
        assert (got == want).all(), (n, np.argwhere(got != want)[:3])


@pytest.mark.parametrize("n", [2, 6, 12, 32])
def test_lsa_kernel_large_batch_vs_oracle_and_properties(n):
    from oracle import gsm_oracle as O
    rng = np.random.default_rng(n)
    B = 4096
    cost = rng.random((B, n, n))
    cost[::2] = rng.integers(0, 3, (B // 2, n, n))           # tie-heavy half
    got = _gsm_lsa(cost)
    assert (np.sort(got, 1) == np.arange(n)).all()           # permutations
    assert (got[:512] == O.lsa(cost[:512])).all()            # oracle, ties included
    # optimality property on the rest: never worse than the identity or a random permutation
    val = np.take_along_axis(cost, got[:, :, None], 2).sum((1, 2))
    assert (val <= np.trace(cost, axis1=1, axis2=2) + 1e-12).all()
    c32 = cost.astype(np.float32)
    g32 = _gsm_lsa(c32)
    v32 = np.take_along_axis(c32, g32[:, :, None], 2).sum((1, 2))
    np.testing.assert_allclose(v32, val, rtol=1e-5)


# ---- API-level equivalences --------------------------------------------------------------
def test_rollout_graph_equals_sequential_steps_and_host_path():
    from gs_marl_b200.env_wrappers import GraphVecEnv
    cfg = make_cfg("polygon", 6, "f64")
    B, T = 96, 7
    o = _squeezed_start(cfg, B, 3)
    rng = np.random.default_rng(0)
    acts = random_actions(cfg, rng, (T, B))
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    seq = []
    for t in range(T):
        env.step(acts[t])
        seq.append({k: v.clone() for k, v in env.buf.items()})
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    roll = env.rollout(acts)
    roll2 = None
    for k in OUT_KEYS:
        assert torch.equal(roll[k], torch.stack([s[k] for s in seq])), k
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    roll2 = env.rollout(acts, out=roll)                        # graph cache hit, same buffers
    assert torch.equal(roll2["obs"][-1], seq[-1]["obs"])
    vec = GraphVecEnv(cfg, B)
    vec.set_state(o.agent_state, o.landmark_pos, o.step_count)
    for t in range(T):
        obs, graph, rew, cost, done, infos = vec.step(acts[t])
    for k in OUT_KEYS:
        assert (vec.buf[k] == _np(seq[-1])[k].view(vec.buf[k].dtype)).all(), k
    assert vec.kernel_launches == T + (T - 1)              # T steps; every step after the dense first one: 1 export kernel
    env.close(); vec.close()


def test_auto_reset_and_done():
    cfg = make_cfg("navigation", 3, "f64", episode_length=3)
    env = _env(cfg, 64, auto_reset=True, seed=9)
    env.reset()
    rng = np.random.default_rng(0)
    for t in range(1, 8):
        obs, graph, rew, cost, done, infos = env.step(random_actions(cfg, rng, (64,)))
        assert bool(done.all()) == (t % 3 == 0)
        ag, lm, tt = env.get_state()
        assert (tt.cpu().numpy() == (t % 3)).all()
        if t % 3 == 0:
            assert (ag[..., 2:] == 0).all()                    # fresh episode: zero velocity
            assert torch.equal(obs[..., 2:4], ag[..., :2])     # obs rows are the post-reset ones
    env.close()


def test_full_size_properties_config1():
    """BASELINE configs[1]: 16384 envs x 3 agents, fp64 verification — checked against the
    oracle in full (the C oracle does this size in well under a second per step)."""
    cfg = make_cfg("navigation", 3, "f64")
    B = 16384
    from oracle import gsm_oracle as O
    O.set_threads(os.cpu_count() or 1)
    o = O.OracleEnv(cfg, B)
    o.reset(1)
    env = _env(cfg, B, seed=1)
    env.reset()
    rng = np.random.default_rng(2)
    for t in range(25):
        a = random_actions(cfg, rng, (B,))
        want = o.step(a)
        env.step(a)
    got = _np(env.buf)
    assert_match(got, want, rtol=F64_RTOL, atol=F64_ATOL, ctx="16384x3 t=25")
    # size-independent properties: adjacency symmetric among agents, cost symmetric-sum even
    adj = got["adj"][..., 0]
    for i in range(3):
        for j in range(3):
            if i != j:
                assert (((adj[:, i] >> j) & 1) == ((adj[:, j] >> i) & 1)).all()
    assert got["cost"].sum() >= 0 and (got["nbr_cnt"] == np.minimum(8, [[bin(int(w)).count("1") for w in r] for r in adj])).all()
    O.set_threads(1)
    env.close()


@pytest.mark.parametrize("name,N,B,kw", [
    ("navigation", 24, 4096, {"max_nbrs": 32}), ("navigation", 48, 4096, {"max_nbrs": 32}),
    ("navigation", 96, 4096, {"max_nbrs": 32}),                      # configs[2]
    ("polygon", 6, 16384, {}), ("polygon", 12, 16384, {}), ("line", 6, 16384, {}), ("line", 12, 16384, {}),   # configs[3]
    ("navigation", 12, 8192, {}),                                    # configs[4], one GPU's share of 65536
])
def test_full_size_other_baseline_configs(name, N, B, kw):
    """BASELINE configs[2..4] at their full per-GPU batch, fp64 verification: one fused 2-step launch with
    the episode end in between (in-kernel re-draw) against the oracle in full, every output."""
    from oracle import gsm_oracle as O
    cfg = make_cfg(name, N, "f64", episode_length=2, **kw)
    O.set_threads(os.cpu_count() or 1)
    try:
        o = O.OracleEnv(cfg, B)
        o.reset(9)
        o.step_count[:] = np.arange(B) % 2                          # half of the envs re-draw after step 0
        env = _env(cfg, B, seed=9)
        env.reset()
        env.set_state(None, None, o.step_count)
        T = 2
        acts = random_actions(cfg, np.random.default_rng(N), (T, B))
        roll = env.rollout(acts, auto_reset=True)
        for t in range(T):
            want = o.step(acts[t])
            got = _np({k: roll[k][t] for k in OUT_KEYS})
            assert_match(got, want, rtol=F64_RTOL, atol=F64_ATOL, ctx=f"{name}{N}x{B} t={t}")
            if want["done"].any():
                o.reset(9, want["done"][:, 0].copy())
        del roll
        env.close()
    finally:
        O.set_threads(1)


# ---- edge cases through the raw C ABI ------------------------------------------------------
def test_out_of_range_actions_mean_no_control():
    """SPEC §2: a discrete index outside the table is u = 0 (both kernels, both precisions)."""
    from oracle import gsm_oracle as O
    for name, N in (("navigation", 3), ("navigation", 24)):
        cfg = make_cfg(name, N, "f64")
        B = 19
        o = _squeezed_start(cfg, B, 4)
        env = _env(cfg, B)
        env.set_state(o.agent_state, o.landmark_pos, o.step_count)
        a = np.random.default_rng(0).integers(-3, 9, (B, N)).astype(np.int32)    # many invalid
        want = o.step(a)
        env.step(a)
        assert_match(_np(env.buf), want, rtol=F64_RTOL, atol=F64_ATOL, ctx=f"{name}{N} bad actions")
        env.close()


@pytest.mark.parametrize("name,N", [("navigation", 3), ("polygon", 6), ("navigation", 24)])
def test_null_outputs_are_skipped_and_single_env(name, N):
    """Any output pointer may be NULL; n_envs = 1; raw ctypes call without the Python mirror."""
    from oracle import gsm_oracle as O
    lib = abi.load_library()
    cfg = make_cfg(name, N, "f64")
    o = _squeezed_start(cfg, 1, 8)
    c, keep = cfg.to_c()
    h = C.c_void_p()
    assert lib.gsm_create(C.byref(c), 1, 0, 0, C.byref(h)) == 0
    ag = torch.as_tensor(o.agent_state).cuda()
    lm = torch.as_tensor(o.landmark_pos).cuda()
    tt = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.gsm_set_state(h, ag.data_ptr(), lm.data_ptr(), tt.data_ptr(), st) == 0
    a_np = random_actions(cfg, np.random.default_rng(1), (1,))
    a = torch.as_tensor(a_np).cuda()
    obs = torch.zeros((1, N, 6), dtype=torch.float64, device="cuda")
    cost = torch.full((1, N), -5.0, dtype=torch.float64, device="cuda")
    io = abi.GsmStepIO()
    io.actions, io.obs, io.cost = a.data_ptr(), obs.data_ptr(), cost.data_ptr()   # everything else NULL
    assert lib.gsm_step(h, C.byref(io), st) == 0, lib.gsm_last_error(h)
    torch.cuda.synchronize()
    want = o.step(a_np)
    np.testing.assert_allclose(obs.cpu().numpy(), want["obs"], rtol=F64_RTOL, atol=F64_ATOL)
    assert (cost.cpu().numpy() == want["cost"]).all()
    io2 = abi.GsmStepIO()                                  # actions missing -> error, not a crash
    assert lib.gsm_step(h, C.byref(io2), st) == -1
    assert b"actions" in lib.gsm_last_error(h)
    assert lib.gsm_destroy(h) == 0


def test_host_path_with_caller_owned_buffers_and_lsa_zero_problems():
    """gsm_step_host into plain (pageable) numpy buffers, not the pinned arena."""
    from oracle import gsm_oracle as O
    lib = abi.load_library()
    cfg = make_cfg("line", 5, "f64")
    B = 21
    o = _squeezed_start(cfg, B, 2)
    c, keep = cfg.to_c()
    h = C.c_void_p()
    assert lib.gsm_create(C.byref(c), B, 0, 0, C.byref(h)) == 0
    assert lib.gsm_set_state_host(h, o.agent_state.ctypes.data, o.landmark_pos.ctypes.data,
                                  o.step_count.ctypes.data) == 0
    bufs = {k: np.zeros(s, d) for k, (d, s) in cfg.io_shapes(B).items()}
    bufs["actions"][...] = random_actions(cfg, np.random.default_rng(3), (B,))
    io = abi.GsmStepIO()
    for k in abi.GsmStepIO.FIELDS:
        setattr(io, k, bufs[k].ctypes.data)
    assert lib.gsm_step_host(h, C.byref(io)) == 0, lib.gsm_last_error(h)
    want = o.step(bufs["actions"])
    assert_match(bufs, want, rtol=F64_RTOL, atol=F64_ATOL, ctx="host caller-owned")
    sz = abi.GsmIoSizes()
    assert lib.gsm_get_io_sizes(h, C.byref(sz)) == 0
    assert sz.nbr_feat == bufs["nbr_feat"].nbytes and sz.adj_words == 1 and sz.real_bytes == 8
    assert lib.gsm_destroy(h) == 0
    assert lib.gsm_lsa(bufs["obs"].ctypes.data, bufs["assign"].ctypes.data, 0, 4, abi.GSM_F64, 0, None) == 0
    assert lib.gsm_lsa(None, None, 1, 4, abi.GSM_F64, 0, None) == -1


@pytest.mark.parametrize("name,N,kw,env_vars", [
    ("navigation", 3, {}, {}),                                   # specialised kernel, in-kernel re-draw
    ("polygon", 6, {}, {}),                                      # specialised + LSA slots re-derived
    ("line", 4, {}, {}),
    ("navigation", 12, {}, {}),                                  # lane kernel, 2 envs per warp
    ("navigation", 40, {"max_nbrs": 16}, {}),                    # lane kernel, env spans 2 warps
    ("navigation", 24, {"max_nbrs": 16}, {"GSM_NO_LANE": "1"}),  # CTA-per-env kernel -> graph fallback
    ("navigation", 33, {"n_obstacles": 0, "max_nbrs": 65}, {"GSM_NO_LANE": "1"}),   # generic -> graph fallback
])
def test_rollout_auto_reset_matches_oracle_resets(name, N, kw, env_vars, monkeypatch):
    """gsm_set_auto_reset + gsm_rollout: envs that finish inside the rollout keep their terminal
    outputs and restart from the SPEC §8 draw — bit-identical to resetting the oracle by hand,
    including envs that finish on different steps."""
    from oracle import gsm_oracle as O
    for k, v in env_vars.items():
        monkeypatch.setenv(k, v)
    cfg = make_cfg(name, N, "f64", episode_length=4, **kw)
    B, T, seed = 13, 11, 99
    o = O.OracleEnv(cfg, B)
    o.reset(seed)
    o.step_count[:] = np.arange(B) % 4                     # staggered episode ends
    env = _env(cfg, B, seed=seed)
    env.reset()
    env.set_state(None, None, o.step_count)
    acts = random_actions(cfg, np.random.default_rng(1), (T, B))
    wants = []
    for t in range(T):
        w = {k: v.copy() for k, v in o.step(acts[t]).items()}
        wants.append(w)
        if w["done"].any():
            o.reset(seed, w["done"][:, 0].copy())
    roll = _np(env.rollout(acts, auto_reset=True))
    for t in range(T):
        assert_match({k: roll[k][t] for k in OUT_KEYS}, wants[t], rtol=F64_RTOL, atol=F64_ATOL,
                     ctx=f"{name}{N} auto-reset t={t}")
    ag, lm, tt = env.get_state()
    np.testing.assert_allclose(ag.cpu().numpy(), o.agent_state, rtol=F64_RTOL, atol=F64_ATOL)
    assert (lm.cpu().numpy() == o.landmark_pos).all() and (tt.cpu().numpy() == o.step_count).all()
    o.reset(seed)                                          # the episode counters advanced identically
    env.reset()
    assert (env.get_state()[0].cpu().numpy() == o.agent_state).all()
    env.close()


def test_fixed_size_env_variants_and_spaces():
    """MultiAgentEnv / MultiAgentConstrainEnv (reference readme.md:29-37): fixed-size views over
    the same kernel; the graph env's per-agent space descriptors."""
    from gs_marl_b200.environment import (MultiAgentEnv, MultiAgentConstrainEnv,
                                          MultiAgentGraphConstrainEnv)
    cfg = make_cfg("navigation", 3, "f64")
    B = 8
    o = _squeezed_start(cfg, B, 12)
    a = random_actions(cfg, np.random.default_rng(0), (B,))
    want = o.step(a)
    flat_want = np.concatenate([want["obs"], want["nbr_feat"].reshape(B, 3, -1)], -1)
    for cls in (MultiAgentConstrainEnv, MultiAgentEnv):
        env = cls(cfg, B)
        start = _squeezed_start(cfg, B, 12)
        env.set_state(start.agent_state, start.landmark_pos, start.step_count)
        out = env.step(a)
        assert len(out) == (5 if cls is MultiAgentConstrainEnv else 4)
        obs, rew = out[0], out[1]
        assert tuple(obs.shape) == (B, 3, 6 + cfg.max_nbrs * 6)
        np.testing.assert_allclose(obs.cpu().numpy(), flat_want, rtol=F64_RTOL, atol=F64_ATOL)
        np.testing.assert_allclose(rew.cpu().numpy(), want["reward"], rtol=F64_RTOL, atol=F64_ATOL)
        if cls is MultiAgentConstrainEnv:
            assert (out[2].cpu().numpy() == want["cost"]).all()
        assert tuple(env.reset().shape) == (B, 3, 6 + cfg.max_nbrs * 6)
        env.close()
    g = MultiAgentGraphConstrainEnv(cfg, B)
    assert g.n == 3 and len(g.observation_space) == 3 and g.observation_space[0].shape == (6,)
    assert g.node_observation_space[0].shape == (cfg.max_nbrs, 6) and g.action_space[0].n == 5
    assert g.share_observation_space[0].shape == (18,)
    g.close()
    c = MultiAgentGraphConstrainEnv(make_cfg("navigation", 3, "f32", action_mode="continuous"), 4)
    assert c.action_space[0].shape == (2,)
    with pytest.raises(ValueError):
        c.step(np.zeros((4, 3), np.float32))               # wrong action shape
    c.close()


def test_team_kernel_enumeration_equals_scipy_order_instance_f32(monkeypatch):
    """fp32 polygon-6: the default instance solves the assignment by exhaustive enumeration and takes the result
    only under a uniqueness certificate (else scipy's procedure, lsa_group2); the G = 3 instance always runs
    scipy's procedure.  Same states -> every output bit-identical, including envs that start on exact ties
    (agents halfway between slots, agents on the slots in reversed order) and 25 steps of random motion, at the
    full BASELINE batch (409 600 assignment problems, 0.35 % of them through the fallback)."""
    from oracle import gsm_oracle as O
    cfg = make_cfg("polygon", 6, "f32")
    B, T, N = 16384, 25, 6
    o = O.OracleEnv(cfg, B)
    o.reset(23)
    ang = (2 * np.arange(N) + 1) * np.pi / N
    ring = cfg.polygon_radius * np.stack([np.cos(ang), np.sin(ang)], -1)
    o.agent_state[:40, :, :2] = o.landmark_pos[:40, :1, :] + ring[None]                      # two equally good shifts
    o.agent_state[40:60, :, :2] = o.landmark_pos[40:60, :1, :] + 3.0 * ring[None]            # far ring: near-ties
    sl = o.landmark_pos[60:80, :1, :] + cfg.polygon_radius * np.asarray(cfg.slot_table).reshape(N, 2)[None]
    o.agent_state[60:80, :, :2] = sl[:, ::-1, :]                                             # on the slots, reversed
    o.agent_state[:80, :, 2:] = 0
    acts = random_actions(cfg, np.random.default_rng(4), (T, B))
    acts[0] = 0
    outs = []
    for team_g in (None, "3"):
        if team_g:
            monkeypatch.setenv("GSM_TEAM_G", team_g)
        env = _env(cfg, B)
        env.set_state(o.agent_state, o.landmark_pos, o.step_count)
        outs.append(_np(env.rollout(acts)))
        env.close()
    for k in OUT_KEYS:
        assert outs[0][k].tobytes() == outs[1][k].tobytes(), k
    a = outs[0]["assign"][..., 0] if outs[0]["assign"].ndim == 4 else outs[0]["assign"]
    assert (np.sort(a.reshape(T, B, N), axis=2) == np.arange(N)).all()                       # permutations


@pytest.mark.parametrize("name,N", [("polygon", 3), ("polygon", 4), ("polygon", 5), ("polygon", 6), ("polygon", 12),
                                    ("line", 3), ("line", 4), ("line", 5), ("line", 6), ("line", 12)])
@pytest.mark.parametrize("kw", [{}, {"share_reward": True, "max_nbrs": 2}, {"team_g": 2}, {"team_g": 3}])
def test_team_kernel_group_lsa_f64(name, N, kw, monkeypatch):
    """env_team_kernel (G lanes per env, N/G assignment rows per lane): fused 25 steps and single
    steps against the oracle, plus a TIE-HEAVY start (all agents on one point, then agents exactly
    on the slots in reversed order) where only scipy's scan order decides the permutation."""
    kw = dict(kw)
    if "team_g" in kw:                                      # the non-default lane groupings (N = 6: G = 2, G = 3 of 4 lanes)
        monkeypatch.setenv("GSM_TEAM_G", str(kw.pop("team_g")))
        if N != 6:
            pytest.skip("only N = 6 has a second grouping")
    cfg = make_cfg(name, N, "f64", **kw)
    B, T = 70, 25
    o = _squeezed_start(cfg, B, 17 + N)
    # ambiguous optima without coincident entities: envs 0..9 agents halfway between consecutive
    # slots (two equally good shifts), envs 10..19 agents exactly on the slots in reversed order
    def slots_of(lm):
        if name == "polygon":
            return lm[:, :1, :] + cfg.polygon_radius * np.asarray(cfg.slot_table)[None]
        f = np.asarray(cfg.slot_table)[:, 0][None, :, None]
        return lm[:, :1, :] + f * (lm[:, 1:2, :] - lm[:, :1, :])
    sl = slots_of(o.landmark_pos[:20])
    if name == "polygon":
        ang = (2 * np.arange(N) + 1) * np.pi / N
        o.agent_state[:10, :, :2] = o.landmark_pos[:10, :1, :] + cfg.polygon_radius * np.stack([np.cos(ang), np.sin(ang)], -1)[None]
    else:
        mid = 0.5 * (sl[:10] + np.roll(sl[:10], -1, axis=1))
        mid[:, -1] = sl[:10, -1] + 0.5 * (sl[:10, -1] - sl[:10, -2])
        o.agent_state[:10, :, :2] = mid
    o.agent_state[10:20, :, :2] = sl[10:20, ::-1, :]
    o.agent_state[:20, :, 2:] = 0
    s0 = o.agent_state.copy()
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    acts = random_actions(cfg, np.random.default_rng(N), (T, B))
    acts[0] = 0                                               # first step: nobody moves -> exact ties
    wants = [{k: v.copy() for k, v in o.step(acts[t]).items()} for t in range(T)]
    roll = _np(env.rollout(acts))
    assert env.kernel_launches == 1
    for t in range(T):
        assert_match({k: roll[k][t] for k in OUT_KEYS}, wants[t], rtol=F64_RTOL, atol=F64_ATOL,
                     ctx=f"{name}{N} team fused t={t}")
    env.set_state(s0, o.landmark_pos, np.zeros(B, np.int32))
    for t in range(2):
        env.step(acts[t])
        assert_match(_np(env.buf), wants[t], rtol=F64_RTOL, atol=F64_ATOL, ctx=f"{name}{N} team step t={t}")
    env.close()


def test_rollout_buffer_collect_is_zero_copy_and_matches_step():
    """collect() with the kernel writing into the rollout buffer's slots gives
    exactly what step()-then-copy gives, and what the oracle gives."""
    from gs_marl_b200.rollout import GraphRolloutBuffer, collect
    cfg = make_cfg("polygon", 4, "f64", episode_length=50)
    B, T = 33, 9
    o = _squeezed_start(cfg, B, 21)
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    buf = GraphRolloutBuffer(env, T)
    env.observe()
    for k in ("obs", "nbr_idx", "nbr_feat", "nbr_cnt", "adj", "assign"):
        buf[k][0].copy_(env.buf[k])

    def policy(obs, graph):                        # deterministic, depends on what it is shown
        s = (obs[..., 2] * 7.0 + graph["nbr_cnt"].to(obs.dtype)).floor().to(torch.int64)
        return (s % 5).to(torch.int32)

    launches0 = env.kernel_launches
    collect(env, policy, buf)
    assert env.kernel_launches - launches0 == T
    # same policy against the oracle
    want0 = o.observe()
    obs_t = torch.as_tensor(want0["obs"]).cuda()
    cnt_t = torch.as_tensor(want0["nbr_cnt"]).cuda()
    for t in range(T):
        a = policy(obs_t, {"nbr_cnt": cnt_t}).cpu().numpy()
        assert (buf["actions"][t].cpu().numpy() == a).all(), t
        w = o.step(a)
        got = {k: buf[k][t + 1].cpu().numpy() for k in ("obs", "nbr_idx", "nbr_feat", "nbr_cnt", "assign")}
        got["adj"] = buf["adj"][t + 1].cpu().numpy().view(np.uint32)
        got.update({k: buf[k][t].cpu().numpy() for k in ("reward", "cost", "done")})
        assert_match(got, w, rtol=F64_RTOL, atol=F64_ATOL, ctx=f"collect t={t}")
        obs_t, cnt_t = buf["obs"][t + 1], buf["nbr_cnt"][t + 1]
    buf.after_update()
    assert torch.equal(buf["obs"][0], buf["obs"][T]) and buf.step == 0
    buf.reset_env()
    assert torch.equal(buf["obs"][0][..., 2:4], env.get_state()[0][..., :2])
    env.close()


@pytest.mark.parametrize("name,N,kw,env_vars", [
    ("navigation", 3, {}, {}), ("navigation", 3, {}, {"GSM_SPEC_P": "8"}), ("navigation", 3, {}, {"GSM_SPEC_P": "1"}),
    ("navigation", 6, {}, {}), ("navigation", 12, {"max_nbrs": 7}, {}), ("navigation", 40, {"max_nbrs": 16}, {}),
    ("navigation", 24, {"max_nbrs": 16}, {"GSM_NO_LANE": "1"}),
    ("navigation", 33, {"n_obstacles": 0, "max_nbrs": 65}, {"GSM_NO_LANE": "1"}),
    ("polygon", 6, {}, {}), ("polygon", 12, {"max_nbrs": 5}, {}), ("line", 5, {}, {}),
    ("polygon", 6, {}, {"GSM_SPEC_P": "1"}), ("line", 4, {}, {"GSM_NO_SPEC": "1"}),
])
@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_no_kernel_writes_outside_its_buffers(name, N, kw, env_vars, dtype, monkeypatch):
    """compute-sanitizer is closed on this pool, so: every output tensor of a fused rollout, a
    single step and a masked reset sits between guard bands that must stay untouched, with an
    env count that leaves ragged tail warps / CTAs."""
    for k, v in env_vars.items():
        monkeypatch.setenv(k, v)
    cfg = make_cfg(name, N, dtype, episode_length=3, **kw)
    B, T, G = 37, 5, 4096                                   # G guard bytes on each side
    env = _env(cfg, B, seed=3)
    env.reset()
    raw, out = {}, {}
    for k in env.OUTPUTS:
        dt, shape = env._shapes[k]
        tdt = env.buf[k].dtype
        n = T * int(np.prod(shape)) * env.buf[k].element_size()
        raw[k] = torch.full((n + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
        out[k] = raw[k][G:G + n].view(tdt).view((T,) + tuple(shape))
    acts = torch.as_tensor(random_actions(cfg, np.random.default_rng(0), (T, B))).cuda()
    env.rollout(acts, out=out, auto_reset=True)
    io = env._make_io({k: v[0] for k, v in out.items()}, acts[0].contiguous())
    env._check(env.lib.gsm_step(env._h, C.byref(io), env._stream()))
    mask = torch.as_tensor(np.arange(B) % 3 == 0).to(torch.uint8).cuda()
    env._check(env.lib.gsm_reset(env._h, 3, mask.data_ptr(), 1, C.byref(io), env._stream()))
    torch.cuda.synchronize()
    for k, r in raw.items():
        assert bool((r[:G] == 0xA5).all()) and bool((r[-G:] == 0xA5).all()), f"{k}: guard band overwritten"
        if k in ("obs", "reward"):
            assert torch.isfinite(out[k].double()).all(), k
    assert int(out["nbr_cnt"].min()) >= 0 and int(out["nbr_cnt"].max()) <= cfg.max_nbrs
    env.close()


@pytest.mark.parametrize("name,N,B,kw", [("navigation", 3, 16384, {}), ("polygon", 12, 16384, {}),
                                         ("line", 6, 16384, {}), ("navigation", 96, 4096, {"max_nbrs": 32}),
                                         ("navigation", 12, 8192, {})])
def test_translation_invariance_at_full_size(name, N, B, kw):
    """Size-independent property at BASELINE's full env counts (no oracle needed): shifting every
    entity of every env by one vector changes nothing but the absolute positions in obs —
    neighbour lists, adjacency, costs, dones and assignments bit-identical (the shift is a
    power of two and positions are pre-rounded so that every difference is exact), rewards and
    relative features identical."""
    cfg = make_cfg(name, N, "f64", **kw)
    from oracle import gsm_oracle as O
    o = O.OracleEnv(cfg, B)
    o.reset(77)
    q = 2.0 ** -20                                       # positions on a grid: shifted sums stay exact
    ag = np.round(o.agent_state * 0.6 / q) * q
    lm = np.round(o.landmark_pos * 0.6 / q) * q
    ag[..., 2:] = 0
    shift = np.array([8.0, -4.0])
    acts = random_actions(cfg, np.random.default_rng(5), (B,))
    outs = []
    for s in (np.zeros(2), shift):
        env = _env(cfg, B)
        a2 = ag.copy(); a2[..., :2] += s
        env.set_state(a2, lm + s, np.zeros(B, np.int32))
        env.observe()
        outs.append(_np({k: v.clone() for k, v in env.buf.items()}))
        env.close()
    a, b = outs
    for k in ("nbr_idx", "nbr_cnt", "adj", "assign"):
        assert (a[k] == b[k]).all(), k
    assert (a["nbr_feat"] == b["nbr_feat"]).all()
    assert (a["obs"][..., :2] == b["obs"][..., :2]).all()
    if name == "navigation":
        assert (a["obs"][..., 4:] == b["obs"][..., 4:]).all()
    else:       # slot = marker + R * unit is rounded after the shift: equal to an ulp of the shifted value
        np.testing.assert_allclose(a["obs"][..., 4:], b["obs"][..., 4:], rtol=0, atol=4e-15)
    assert (b["obs"][..., 2:4] - a["obs"][..., 2:4] == shift).all()
    assert a["nbr_cnt"].sum() > 0


WIDE_VARIANTS = [{}, {"share_reward": True}, {"own_goal_always": False}, {"cost_obstacles": False},
                 {"action_mode": "continuous"}, {"max_nbrs": 4}, {"max_nbrs": 5, "share_reward": True},
                 {"sensing_radius": 0.4}, {"sensing_radius": 3.0, "max_nbrs": 8}]


@pytest.mark.parametrize("kw", WIDE_VARIANTS)
@pytest.mark.parametrize("B", [1, 3, 130, 257])
@pytest.mark.parametrize("N", [3, 4, 5])
def test_wide_kernel_variants_f64(kw, B, N):
    """env_wide_kernel (navigation-3: one lane per other entity, shared-memory entity table, staged outputs,
    bulk copies) on every scenario switch, on max_nbrs below / at the number of others (the run-time-K
    instance and the K = 8 instance), and on ragged batches (warps with 1..3 of 4 envs take the word-wise
    copy path): fused 25 steps with in-kernel auto-reset, single steps and observe against the oracle.
    N = 4 (11 others on 16 lanes, 2 envs per warp) exercises the instance whose groups have idle lanes, N = 5
    the 8-byte scalar pieces (fp32) and the word-wise copy path for every warp (fp64: 35 pieces > 32 lanes)."""
    from oracle import gsm_oracle as O
    cfg = make_cfg("navigation", N, "f64", episode_length=7, **kw)
    seed, T = 5 + B, 25
    o = O.OracleEnv(cfg, B)
    o.reset(seed)
    o.agent_state[..., :2] *= 0.45                             # squeezed: contacts from the first step on
    o.landmark_pos *= 0.45
    o.step_count[:] = np.arange(B) % 7                         # staggered episode ends
    env = _env(cfg, B, seed=seed)
    env.reset()
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    obs, graph = env.observe()
    assert_match({**_np(graph), "obs": obs.cpu().numpy()}, o.observe(), rtol=F64_RTOL, atol=F64_ATOL,
                 ctx=f"wide {kw} observe", keys=("obs", "nbr_idx", "nbr_feat", "nbr_cnt", "adj"))
    acts = random_actions(cfg, np.random.default_rng(B), (T, B))
    launches0 = env.kernel_launches
    roll = _np(env.rollout(acts, auto_reset=True))
    assert env.kernel_launches - launches0 == 1                 # ONE fused launch
    total_cost = 0.0
    for t in range(T):
        want = {k: v.copy() for k, v in o.step(acts[t]).items()}
        assert_match({k: roll[k][t] for k in OUT_KEYS}, want, rtol=F64_RTOL, atol=F64_ATOL, ctx=f"wide {kw} B={B} t={t}")
        total_cost += want["cost"].sum()
        if want["done"].any():
            o.reset(seed, want["done"][:, 0].copy())
    ag, lm, tt = env.get_state()
    np.testing.assert_allclose(ag.cpu().numpy(), o.agent_state, rtol=F64_RTOL, atol=F64_ATOL)
    assert (lm.cpu().numpy() == o.landmark_pos).all() and (tt.cpu().numpy() == o.step_count).all()
    if B >= 130:
        assert total_cost > 0, "case never exercised a collision"
    # single steps after the rollout (the one-step instance of the same kernel)
    for t in range(3):
        a1 = random_actions(cfg, np.random.default_rng(100 + t), (B,))
        want = o.step(a1)
        env.step(a1)
        assert_match(_np(env.buf), want, rtol=F64_RTOL, atol=F64_ATOL, ctx=f"wide {kw} single step {t}")
    env.close()


@pytest.mark.parametrize("N", [4, 5])
def test_wide_kernel_f32_small_teams_vs_oracle(N):
    """fp32 instances of the wide kernel for N = 4 / 5 (16-byte / 8-byte scalar pieces, bulk feature copies):
    one fused launch against the fp32 oracle on rows away from contacts and thresholds."""
    from oracle import gsm_oracle as O
    cfg = make_cfg("navigation", N, "f32")
    B, T = 258, 6
    o = O.OracleEnv(cfg, B)
    o.reset(3)
    env = _env(cfg, B)
    env.set_state(o.agent_state, o.landmark_pos, o.step_count)
    acts = random_actions(cfg, np.random.default_rng(N), (T, B))
    out = _np(env.rollout(acts))
    touched, _ = _contact_touched(cfg, o.agent_state, o.landmark_pos, 0.03)
    clean = ~touched
    checked = 0
    for t in range(T):
        want = o.step(acts[t])
        touched, d_aa = _contact_touched(cfg, o.agent_state, o.landmark_pos, 0.03)
        clean &= ~touched
        dirty_near = ((d_aa < cfg.sensing_radius + 0.5) & ~clean[:, None, :]).any(2)
        ok = clean & ~dirty_near & ~near_threshold_rows(cfg, o.agent_state, o.landmark_pos, 1e-4)
        for k in ("nbr_cnt", "cost", "nbr_idx", "adj", "done", "assign"):
            assert (out[k][t][ok] == want[k][ok]).all(), (N, t, k)
        for k in ("obs", "nbr_feat", "reward"):
            np.testing.assert_allclose(out[k][t][ok], want[k][ok], rtol=1e-4, atol=2e-5, err_msg=f"nav{N} t={t} {k}")
        checked += int(ok.sum())
    assert checked > 0.05 * T * B * N
    env.close()


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_wide_kernel_equals_spec_kernel_and_unaligned_buffers(dtype, monkeypatch):
    """The wide kernel against the kernel it replaced (GSM_NO_WIDE=1 -> env_steps_kernel) on the same states
    and actions: integer outputs identical, reals to rounding (the pair forces are summed in another order);
    then the same rollout into output tensors that start 4 / 8 bytes off a 16-byte boundary (no bulk copies:
    the word-wise path) — bit-identical to the aligned run."""
    cfg = make_cfg("navigation", 3, dtype, episode_length=9)
    B, T = 515, 20
    rng = np.random.default_rng(1)
    acts = random_actions(cfg, rng, (T, B))
    a = _env(cfg, B, seed=2); a.reset()
    ra = {k: v.clone() for k, v in a.rollout(acts, auto_reset=True).items()}
    a.close()
    # unaligned outputs: every tensor is a view that starts one element into a larger allocation
    c = _env(cfg, B, seed=2); c.reset()
    out = {}
    for k in c.OUTPUTS:
        full = c._alloc(k, (T,))
        big = torch.zeros(full.numel() + 8, dtype=full.dtype, device=full.device)
        out[k] = big[1:1 + full.numel()].view(full.shape)
        assert out[k].data_ptr() % 16 != 0 or out[k].element_size() >= 16
    rb = c.rollout(acts, out=out, auto_reset=True)
    for k in OUT_KEYS:
        assert torch.equal(ra[k], rb[k]), (k, "unaligned buffers")
    c.close()
    monkeypatch.setenv("GSM_NO_WIDE", "1")
    b = _env(cfg, B, seed=2); b.reset()
    rs = b.rollout(acts, auto_reset=True)
    tol = dict(rtol=1e-4, atol=2e-5) if dtype == "f32" else dict(rtol=F64_RTOL, atol=F64_ATOL)
    ga, gs = _np(ra), _np(rs)
    if dtype == "f64":
        assert_match({k: ga[k] for k in OUT_KEYS}, {k: gs[k] for k in OUT_KEYS}, ctx="wide vs spec", **tol)
    else:                                                       # fp32: compare the first steps, before rounding differences grow at contacts
        for k in ("nbr_cnt", "adj", "done", "assign"):
            assert (ga[k][:2] == gs[k][:2]).mean() > 0.999, k
        np.testing.assert_allclose(ga["obs"][:2], gs["obs"][:2], **tol)
    b.close()


@pytest.mark.parametrize("name,N,dtype,B", [("navigation", 3, "f32", 1027), ("navigation", 12, "f32", 130),
                                             ("polygon", 6, "f64", 65), ("navigation", 3, "f64", 33)])
def test_host_sparse_export_is_bit_identical_to_dense_copy(name, N, dtype, B):
    """gsm_set_host_outputs: the sparse export (a kernel writes only the nbr_cnt valid rows of nbr_feat
    into the mapped arena, clears rows that stop being valid) against one dense D2H copy per call,
    over steps, full and masked resets; then with outputs switched off and on again."""
    from gs_marl_b200.env_wrappers import GraphVecEnv
    cfg = make_cfg(name, N, dtype, episode_length=6)
    a, b = GraphVecEnv(cfg, B, seed=3), GraphVecEnv(cfg, B, seed=3)
    b.set_host_outputs(None, sparse=False)
    rng = np.random.default_rng(B)

    def same(keys=OUT_KEYS):
        for k in keys:
            assert a.buf[k].tobytes() == b.buf[k].tobytes(), (name, N, k)
    a.reset(); b.reset()
    same([k for k in OUT_KEYS if k not in ("reward", "cost", "done")])
    for t in range(20):
        acts = random_actions(cfg, rng, (B,))
        a.step(acts); b.step(acts)
        same()
        if t % 6 == 5:
            m = (rng.random(B) < 0.4).astype(np.uint8)
            a.reset(m); b.reset(m)
            same()
        if t == 12:
            a.reset(); b.reset()
            same()
    # rows beyond nbr_cnt really are padding on the host
    cnt = a.buf["nbr_cnt"]
    pad = np.arange(cfg.max_nbrs)[None, None, :] >= cnt[..., None]
    assert (a.buf["nbr_idx"][pad] == -1).all() and (a.buf["nbr_feat"][pad] == 0).all()
    # outputs switched off keep their old host contents, the others still match
    keep_idx = a.buf["nbr_idx"].copy()
    a.set_host_outputs([k for k in OUT_KEYS if k not in ("nbr_idx", "assign")])
    for t in range(3):
        acts = random_actions(cfg, rng, (B,))
        a.step(acts); b.step(acts)
    assert (a.buf["nbr_idx"] == keep_idx).all()
    same([k for k in OUT_KEYS if k not in ("nbr_idx", "assign")])
    a.set_host_outputs(None)                                   # back on: dense re-synchronisation, then sparse again
    for t in range(3):
        acts = random_actions(cfg, rng, (B,))
        a.step(acts); b.step(acts)
        same()
    assert a.kernel_launches > b.kernel_launches               # the export kernels are counted
    a.close(); b.close()


def test_graph_vec_env_host_auto_reset():
    """numpy drop-in with auto_reset: after `done` the returned obs/graph rows are those of the
    freshly drawn episode, reward/cost/done keep the terminal step's values."""
    from gs_marl_b200.env_wrappers import GraphVecEnv
    from oracle import gsm_oracle as O
    cfg = make_cfg("navigation", 3, "f64", episode_length=2)
    B = 17
    vec = GraphVecEnv(cfg, B, auto_reset=True, seed=4)
    o = O.OracleEnv(cfg, B)
    o.reset(4)
    obs, graph = vec.reset()
    assert (obs == o.observe()["obs"]).all()
    rng = np.random.default_rng(0)
    for t in range(1, 5):
        a = random_actions(cfg, rng, (B,))
        want = o.step(a)
        obs, graph, rew, cost, done, infos = vec.step(a)
        np.testing.assert_allclose(rew, want["reward"], rtol=F64_RTOL, atol=F64_ATOL)
        assert (cost == want["cost"]).all() and (done == want["done"]).all()
        if t % 2 == 0:
            assert done.all()
            o.reset(4)
            w0 = o.observe()
            np.testing.assert_allclose(obs, w0["obs"], rtol=F64_RTOL, atol=F64_ATOL)
            assert (graph["nbr_idx"] == w0["nbr_idx"]).all()
        else:
            np.testing.assert_allclose(obs, want["obs"], rtol=F64_RTOL, atol=F64_ATOL)
    ag, lm, tt = vec.get_state()
    assert (tt == 0).all() and (ag == o.agent_state).all()
    vec.close()


@pytest.mark.parametrize("name,N,B", [("navigation", 3, 8192), ("navigation", 12, 2048), ("polygon", 6, 4096)])
def test_fp32_production_mode_is_statistically_equal_to_fp64(name, N, B):
    """200 random-action steps with in-kernel auto-reset: individual fp32 trajectories drift from
    the fp64 ones after contacts (stiff springs), but the rollout statistics a learner sees must
    not: mean reward, mean cost, mean neighbour count, done count."""
    T, E = 200, 25
    stats = {}
    for dtype in ("f64", "f32"):
        cfg = make_cfg(name, N, dtype, episode_length=E)
        env = _env(cfg, B, seed=11)
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        acts = torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.int32)
        roll = env.rollout(acts, auto_reset=True)
        stats[dtype] = dict(reward=roll["reward"].double().mean().item(), cost=roll["cost"].double().mean().item(),
                            cnt=roll["nbr_cnt"].double().mean().item(), done=int(roll["done"].sum().item()),
                            first_obs=roll["obs"][0].double().cpu().numpy())
        env.close()
    a, b = stats["f64"], stats["f32"]
    np.testing.assert_allclose(b["first_obs"], a["first_obs"], rtol=1e-4, atol=1e-5)   # same start, one step
    assert a["done"] == b["done"] == (T // E) * B * N
    assert abs(a["reward"] - b["reward"]) < 2e-3 * abs(a["reward"]), (a["reward"], b["reward"])
    assert abs(a["cnt"] - b["cnt"]) < 2e-3 * a["cnt"], (a["cnt"], b["cnt"])
    assert a["cost"] > 0 and abs(a["cost"] - b["cost"]) < 0.03 * a["cost"], (a["cost"], b["cost"])


# ---- intra-GPU sub-shards on concurrent streams -----------------------------------------------
@pytest.mark.parametrize("name,N,kw,env_vars", [
    ("navigation", 3, {}, {}),                                   # specialised kernel
    ("navigation", 12, {}, {}),                                  # lane kernel
    ("polygon", 6, {}, {}),                                      # team kernel
    ("navigation", 33, {"n_obstacles": 0, "max_nbrs": 65}, {"GSM_NO_LANE": "1"}),   # generic -> CUDA-graph rollout
])
@pytest.mark.parametrize("dtype,S", [("f64", 3), ("f32", 4)])
def test_stream_sharded_env_is_bit_identical_to_one_handle(name, N, kw, env_vars, dtype, S, monkeypatch):
    """StreamShardedEnv (S handles on S streams filling ONE set of [T][n_envs] buffers through
    gsm_set_slot_envs) == one handle over the same envs, bit for bit: reset, masked reset, step,
    fused rollouts with in-kernel auto-reset, back-to-back un-joined rollouts, final state.  The
    fp64 leg is also checked against the oracle."""
    from gs_marl_b200.environment import StreamShardedEnv
    for k, v in env_vars.items():
        monkeypatch.setenv(k, v)
    cfg = make_cfg(name, N, dtype, episode_length=5, **kw)
    B, T, seed = 29, 7, 5                                        # ragged: 29 envs over 3 / 4 shards
    one = _env(cfg, B, seed=seed, env_offset=100)
    many = StreamShardedEnv(cfg, B, n_streams=S, seed=seed, env_offset=100)
    assert [hi - lo for lo, hi in many.bounds] == sorted([hi - lo for lo, hi in many.bounds], reverse=True)
    assert sum(hi - lo for lo, hi in many.bounds) == B

    def same(a, b, ctx):
        a, b = _np(a), _np(b)
        for k in a:
            if k in b:
                assert np.array_equal(a[k], b[k], equal_nan=True), (ctx, k)

    o1, g1 = one.reset()
    o2, g2 = many.reset()
    same(dict(obs=o1, **g1), dict(obs=o2, **g2), "reset")
    rng = np.random.default_rng(3)
    acts = random_actions(cfg, rng, (3 * T, B))
    r1 = one.step(acts[0]); r2 = many.step(acts[0])
    first_step = _np(dict(obs=r2[0], reward=r2[2], cost=r2[3], done=r2[4], assign=r2[5]["assign"], **r2[1]))
    same(dict(obs=r1[0], reward=r1[2], cost=r1[3], done=r1[4], **r1[1]), first_step, "step")
    mask = torch.tensor((np.arange(B) % 3 == 0).astype(np.uint8), device="cuda")
    o1, g1 = one.reset(mask); o2, g2 = many.reset(mask)
    same(dict(obs=o1, **g1), dict(obs=o2, **g2), "masked reset")
    # three back-to-back rollouts, the sharded ones left un-joined until the end
    outs1 = [one.rollout(acts[q * T:(q + 1) * T], auto_reset=True) for q in range(3)]
    outs2 = [many.rollout(acts[q * T:(q + 1) * T], auto_reset=True, join=False) for q in range(3)]
    many.join()
    torch.cuda.synchronize()
    for q in range(3):
        same(outs1[q], outs2[q], f"rollout {q}")
    for a, b in zip(one.get_state(), many.get_state()):
        assert torch.equal(a, b)
    assert many.kernel_launches >= S
    if dtype == "f64":
        from oracle import gsm_oracle as O
        o = O.OracleEnv(cfg, B, env_offset=100)
        o.reset(seed)
        w = {k: v.copy() for k, v in o.step(acts[0]).items()}
        assert_match(first_step, w, rtol=F64_RTOL, atol=F64_ATOL, ctx="sharded step vs oracle")
    one.close(); many.close()


def test_slot_envs_argument_checks():
    cfg = make_cfg("navigation", 3, "f32")
    env = _env(cfg, 8)
    assert env.lib.gsm_set_slot_envs(env._h, 4) == -1            # narrower than the handle
    assert b"slot_envs" in env.lib.gsm_last_error(env._h)
    assert env.lib.gsm_set_slot_envs(env._h, 8) == 0 and env.lib.gsm_set_slot_envs(env._h, 0) == 0
    assert env.lib.gsm_set_slot_envs(None, 8) == -1
    env.close()


def test_checkpoint_resume_with_episode_counters():
    """A checkpoint = agent_state + landmark_pos + step_count + EPISODE counters: restored into a fresh
    handle, the following steps — including the re-draws of envs that finish — repeat bit for bit.
    Without the episode counters the re-draws would come from episode 1 again."""
    cfg = make_cfg("navigation", 3, "f32").replace(episode_length=4)
    a = _env(cfg, 200, seed=6, env_offset=11, auto_reset=True)
    a.reset()
    rng = np.random.default_rng(100)
    acts = [torch.from_numpy(random_actions(cfg, rng, (200,))).cuda() for s in range(14)]
    for s in range(6):                                   # one re-draw behind us, t = 2 inside episode 2
        a.step(acts[s])
    state, ep = a.get_state(), a.get_episode()
    assert int(ep.min()) == int(ep.max()) == 2
    b = _env(cfg, 200, seed=6, env_offset=11, auto_reset=True)
    b.set_state(*state)
    b.set_episode(ep)
    for s in range(6, 14):                               # two more re-draws
        ra, rb = a.step(acts[s]), b.step(acts[s])
    torch.cuda.synchronize()
    for k in OUT_KEYS:
        assert torch.equal(a.buf[k], b.buf[k]), k
    c = _env(cfg, 200, seed=6, env_offset=11, auto_reset=True)      # the same WITHOUT the counters diverges
    c.set_state(*state)
    for s in range(6, 14):
        c.step(acts[s])
    torch.cuda.synchronize()
    assert not torch.equal(a.buf["obs"], c.buf["obs"])
    for e in (a, b, c):
        e.close()
