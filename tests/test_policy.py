"""the declared graph actor (SPEC.md §10) and the fused collect loop.

CPU: the numpy oracle against an independent plain-torch fp32 formulation, numpy Philox against
the C oracle's (Random123-pinned).  GPU: `gsm_policy_act` against the oracle (logits, log-probs,
sampled actions), sharding invariance of the draws, and `gsm_collect` against the step-by-step
Python collect loop driving the same kernels."""
import ctypes as C

import numpy as np
import pytest
import torch

from gs_marl_b200 import abi
from gs_marl_b200.policy import GraphAttentionActor
from oracle import gsm_oracle as O, policy_oracle as P
from tests._util import make_cfg


def synth(R, K, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    obs = (rng.standard_normal((R, 6)) * scale).astype(np.float32)
    cnt = rng.integers(0, K + 1, R).astype(np.int32)
    feat = (rng.standard_normal((R, K, 6)) * scale).astype(np.float32)
    feat[np.arange(K)[None, :] >= cnt[:, None]] = 0          # the env zero-fills padded rows
    return obs, feat, cnt


def test_numpy_philox_matches_c_oracle():
    rng = np.random.default_rng(1)
    c = rng.integers(0, 2 ** 32, (64, 4), dtype=np.uint64)
    k0, k1 = 0xDEADBEEF, 0x12345678
    got = np.stack(P.philox4x32_10(c[:, 0], c[:, 1], c[:, 2], c[:, 3], k0, k1), 1)
    for i in range(64):
        want = O.philox(int(c[i, 0]), int(c[i, 1]), int(c[i, 2]), int(c[i, 3]), k0, k1)
        assert tuple(int(x) for x in got[i]) == tuple(int(x) for x in want)


@pytest.mark.parametrize("n_actions", [5, 9])
def test_oracle_matches_torch_formulation(n_actions):
    actor = GraphAttentionActor(n_actions, seed=3)
    obs, feat, cnt = synth(512, 8, seed=2)
    w = P.weights_from_state_dict(actor.state_dict())
    z = P.logits(w, obs, feat, cnt, np.float64)
    with torch.no_grad():
        zt = actor.logits_autograd(torch.from_numpy(obs), {"nbr_feat": torch.from_numpy(feat),
                                                           "nbr_cnt": torch.from_numpy(cnt)}).numpy()
    np.testing.assert_allclose(z, zt, rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        vt = actor.values_autograd(torch.from_numpy(obs), {"nbr_feat": torch.from_numpy(feat),
                                                           "nbr_cnt": torch.from_numpy(cnt)}).numpy()
    np.testing.assert_allclose(P.values(w, obs, feat, cnt), vt, rtol=1e-4, atol=1e-5)
    # padded rows must not matter: garbage in them leaves the logits unchanged
    junk = feat.copy()
    junk[np.arange(8)[None, :] >= cnt[:, None]] = 1e3
    np.testing.assert_array_equal(P.logits(w, obs, junk, cnt), z)


def test_pack_layout_roundtrip():
    actor = GraphAttentionActor(5, seed=4)
    w = actor.pack()
    assert w.struct_size == C.sizeof(abi.GsmPolicyWeights) == 9528 and w.n_actions == 5
    np.testing.assert_array_equal(np.ctypeslib.as_array(w.ego_w), actor.ego.weight.detach().numpy())
    np.testing.assert_array_equal(np.ctypeslib.as_array(w.head_w)[:5], actor.head.weight.detach().numpy())
    assert np.all(np.ctypeslib.as_array(w.head_w)[5:] == 0)
    assert abs(w.att_b - actor.att.bias.item()) < 1e-7
    np.testing.assert_array_equal(np.ctypeslib.as_array(w.value_w), actor.value.weight.detach().numpy())
    assert C.sizeof(abi.GsmPolicyIO) == 96


def test_policy_bad_arguments_without_device():
    lib = abi.load_library()
    actor = GraphAttentionActor(5)
    w = actor.pack()
    io = abi.GsmPolicyIO()
    assert lib.gsm_policy_act(C.byref(w), C.byref(io), 0, None) == -1        # NULL pointers
    w.struct_size = 4
    assert lib.gsm_policy_act(C.byref(w), C.byref(io), 0, None) == -3
    w.struct_size = C.sizeof(abi.GsmPolicyWeights); w.n_actions = 4
    assert lib.gsm_policy_act(C.byref(w), C.byref(io), 0, None) == -4
    assert b"5 and 9" in lib.gsm_policy_last_error()
    if not torch.cuda.is_available():
        with pytest.raises(abi.GsmError):
            actor.act(torch.zeros(4, 6), {"nbr_feat": torch.zeros(4, 8, 6), "nbr_cnt": torch.zeros(4, dtype=torch.int32)})


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n_actions,R,K", [(5, 49152, 8), (9, 4097, 5), (5, 31, 1), (5, 1, 23)])
def test_actor_kernel_matches_oracle(n_actions, R, K):
    actor = GraphAttentionActor(n_actions, seed=5)
    obs, feat, cnt = synth(R, K, seed=6)
    w = P.weights_from_state_dict(actor.state_dict())
    dev = torch.device("cuda", 0)
    g = {"nbr_feat": torch.from_numpy(feat).to(dev), "nbr_cnt": torch.from_numpy(cnt).to(dev)}
    a, lp, z, v = actor.act(torch.from_numpy(obs).to(dev), g, seed=77, step=3, row_offset=1000, want_logits=True,
                            want_values=True)
    np.testing.assert_allclose(v.cpu().numpy(), P.values(w, obs, feat, cnt), rtol=1e-4, atol=2e-5)
    a_only, lp_only = actor.act(torch.from_numpy(obs).to(dev), g, seed=77, step=3, row_offset=1000)
    assert torch.equal(a_only, a) and torch.equal(lp_only, lp)          # the actor-only instance agrees
    # garbage (NaN) in the padded rows must not reach any output
    junk = feat.copy()
    junk[np.arange(K)[None, :] >= cnt[:, None]] = np.nan
    aj, lpj = actor.act(torch.from_numpy(obs).to(dev), {"nbr_feat": torch.from_numpy(junk).to(dev),
                                                        "nbr_cnt": g["nbr_cnt"]}, seed=77, step=3, row_offset=1000)
    assert torch.equal(aj, a) and torch.equal(lpj, lp)
    a0, lp0, z0, margin = P.act(w, obs, feat, cnt, seed=77, step=3, row_offset=1000)
    # fp32 production tolerance (north_star: 1e-4 relative)
    np.testing.assert_allclose(z.cpu().numpy(), z0, rtol=1e-4, atol=2e-5)
    safe = margin > 1e-3
    assert safe.mean() > 0.99
    np.testing.assert_array_equal(a.cpu().numpy()[safe], a0[safe])
    np.testing.assert_allclose(lp.cpu().numpy()[safe], lp0[safe], rtol=1e-4, atol=2e-5)
    # greedy
    ag, _ = actor.act(torch.from_numpy(obs).to(dev), g, greedy=True)
    a1, _, _, m1 = P.act(w, obs, feat, cnt, greedy=True)
    np.testing.assert_array_equal(ag.cpu().numpy()[m1 > 1e-3], a1[m1 > 1e-3])


@pytest.mark.gpu
def test_actor_draws_are_sharding_invariant_and_step_dependent():
    actor = GraphAttentionActor(5, seed=7)
    obs, feat, cnt = synth(6000, 8, seed=8)
    dev = torch.device("cuda", 0)
    to = lambda x: torch.from_numpy(x).to(dev)
    full, _ = actor.act(to(obs), {"nbr_feat": to(feat), "nbr_cnt": to(cnt)}, seed=9, step=4)
    lo, _ = actor.act(to(obs[:2500]), {"nbr_feat": to(feat[:2500]), "nbr_cnt": to(cnt[:2500])}, seed=9, step=4)
    hi, _ = actor.act(to(obs[2500:]), {"nbr_feat": to(feat[2500:]), "nbr_cnt": to(cnt[2500:])}, seed=9, step=4,
                      row_offset=2500)
    assert torch.equal(full, torch.cat([lo, hi]))
    other, _ = actor.act(to(obs), {"nbr_feat": to(feat), "nbr_cnt": to(cnt)}, seed=9, step=5)
    assert (other != full).float().mean() > 0.2
    # sampled frequencies follow softmax(logits)
    z = P.logits(P.weights_from_state_dict(actor.state_dict()), obs[:1], feat[:1], cnt[:1])[0]
    p = np.exp(z - z.max()); p /= p.sum()
    rep = 200000
    o1 = to(np.repeat(obs[:1], rep, 0)); f1 = to(np.repeat(feat[:1], rep, 0)); c1 = to(np.repeat(cnt[:1], rep, 0))
    s, _ = actor.act(o1, {"nbr_feat": f1, "nbr_cnt": c1}, seed=11, step=0)
    freq = np.bincount(s.cpu().numpy(), minlength=5) / rep
    assert np.abs(freq - p).max() < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("scn,N", [("navigation", 3), ("navigation", 12), ("polygon", 6)])
def test_gsm_collect_equals_stepwise_collect(scn, N):
    """gsm_collect (C loop of actor + env-step launches into the buffer slots) == the Python
    collect loop calling the same kernels one step at a time: bit-exact buffers."""
    from gs_marl_b200.environment import MultiAgentGraphConstrainEnv
    from gs_marl_b200.rollout import GraphRolloutBuffer, collect, collect_fused
    cfg = make_cfg(scn, N, "f32").replace(episode_length=3)     # episodes end (and restart) inside the rollout
    actor = GraphAttentionActor(len(cfg.discrete_u), seed=12)
    T, n_envs = 7, 300
    bufs = []
    for mode in ("python", "fused", "graph"):
        env = MultiAgentGraphConstrainEnv(cfg, n_envs, env_offset=40, seed=5, auto_reset=True)
        buf = GraphRolloutBuffer(env, T)
        buf.reset_env()
        if mode == "python":
            t_box = [0]

            def policy(obs, graph):
                a, lp, v = actor.act(obs, graph, seed=21, step=100 + t_box[0], row_offset=40 * N, want_values=True)
                buf["logp"][t_box[0]].copy_(lp)
                buf["values"][t_box[0]].copy_(v)
                t_box[0] += 1
                return a
            collect(env, policy, buf)
        elif mode == "fused":
            collect_fused(env, actor, buf, seed=21, first_step=100)
        else:
            g = collect_fused(env, actor, buf, seed=21, first_step=100, graph=True)
            for k in ("actions", "reward", "logp", "values"):
                buf[k].zero_()
            g.replay()
        torch.cuda.synchronize()
        bufs.append({k: v.clone() for k, v in buf.data.items()})
        env.close()
    for k in bufs[0]:
        assert torch.equal(bufs[0][k], bufs[1][k]), k
        assert torch.equal(bufs[0][k], bufs[2][k]), k
    assert bufs[0]["actions"].float().std() > 0
    d = bufs[0]["done"].cpu().numpy()
    assert d[2].all() and d[5].all() and not d[[0, 1, 3, 4, 6]].any()      # done every 3rd step: envs restarted


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_collect_fused_over_stream_shards_equals_one_handle(graph):
    """Sub-shards on their own streams (StreamShardedEnv) running their own actor -> env chains fill
    the same buffer with exactly what one handle produces: resets and Gumbel draws are keyed by the
    global env / agent index."""
    from gs_marl_b200.environment import MultiAgentGraphConstrainEnv, StreamShardedEnv
    from gs_marl_b200.rollout import GraphRolloutBuffer, collect_fused
    cfg = make_cfg("navigation", 3, "f32")
    actor = GraphAttentionActor(len(cfg.discrete_u), seed=13)
    T, n_envs = 6, 1000
    outs = []
    for S in (1, 3):
        env = (MultiAgentGraphConstrainEnv(cfg, n_envs, env_offset=7, seed=9) if S == 1 else
               StreamShardedEnv(cfg, n_envs, n_streams=S, env_offset=7, seed=9))
        buf = GraphRolloutBuffer(env, T)
        buf.reset_env()
        r = collect_fused(env, actor, buf, seed=3, first_step=50, graph=graph)
        if graph:
            buf["actions"].zero_(); buf["logp"].zero_(); buf["values"].zero_()
            r.replay()
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in buf.data.items()})
        env.close()
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_gae_oracle_closed_form():
    """gamma = lam = 1, no dones: returns[t] = sum of the rewards from t on + V[T]."""
    rng = np.random.default_rng(0)
    T, R = 9, 5
    rw, cs = rng.standard_normal((T, R)), rng.random((T, R))
    v = rng.standard_normal((T + 1, R, 2))
    ret, adv = P.gae(rw, cs, v, np.zeros((T, R), np.uint8), 1.0, 1.0)
    np.testing.assert_allclose(ret[..., 0], np.cumsum(rw[::-1], 0)[::-1] + v[T, :, 0], atol=1e-12)
    np.testing.assert_allclose(ret[..., 1], np.cumsum(cs[::-1], 0)[::-1] + v[T, :, 1], atol=1e-12)
    np.testing.assert_allclose(adv, ret - v[:-1], atol=1e-12)
    # a done cuts the bootstrap and the accumulation
    d = np.zeros((T, R), np.uint8); d[4] = 1
    ret2, _ = P.gae(rw, cs, v, d, 1.0, 1.0)
    np.testing.assert_allclose(ret2[4, :, 0], rw[4], atol=1e-12)
    np.testing.assert_allclose(ret2[:4], P.gae(rw[:5], cs[:5], v[:6], d[:5], 1.0, 1.0)[0][:4], atol=1e-12)


@pytest.mark.gpu
def test_gae_kernel_matches_oracle_on_a_collected_rollout():
    from gs_marl_b200.environment import MultiAgentGraphConstrainEnv
    from gs_marl_b200.rollout import GraphRolloutBuffer, collect_fused
    cfg = make_cfg("navigation", 3, "f32").replace(episode_length=5)
    actor = GraphAttentionActor(len(cfg.discrete_u), seed=14)
    env = MultiAgentGraphConstrainEnv(cfg, 777, seed=2)
    T = 12                                               # episodes end inside the rollout: done flags set
    buf = GraphRolloutBuffer(env, T)
    buf.reset_env()
    collect_fused(env, actor, buf, seed=1)
    *_, vT = actor.act(buf["obs"][T], buf.graph(T), want_values=True)
    buf["values"][T].copy_(vT)
    ret, adv = buf.compute_returns(0.97, 0.9)
    torch.cuda.synchronize()
    done = buf["done"].cpu().numpy()
    assert done.any() and not done.all()
    r0, a0 = P.gae(buf["reward"].cpu().numpy().reshape(T, -1), buf["cost"].cpu().numpy().reshape(T, -1),
                   buf["values"].cpu().numpy().reshape(T + 1, -1, 2), done.reshape(T, -1), 0.97, 0.9)
    np.testing.assert_allclose(ret.cpu().numpy().reshape(T, -1, 2), r0, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(adv.cpu().numpy().reshape(T, -1, 2), a0, rtol=1e-4, atol=1e-4)
    env.close()


def test_oracle_reproduces_committed_policy_golden():
    """tests/golden/policy_golden.npz (generator: tests/golden/make_policy_golden.py) pins the actor /
    sampling / GAE oracle: the target the CUDA kernels are compared with must not drift."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_golden.npz"))
    w = {k[2:]: g[k] for k in g.files if k.startswith("w_")}
    a, lp, z, margin = P.act(w, g["obs"], g["feat"], g["cnt"], seed=99, step=7, row_offset=12345)
    np.testing.assert_allclose(z, g["logits"], rtol=1e-10, atol=1e-10)
    # float32 log may differ by an ulp between SIMD paths / numpy builds: tolerance, and actions only
    # where the two largest perturbed logits are further apart than that could matter
    np.testing.assert_allclose(P.gumbel(96, 5, 99, 7, 12345), g["gumbel"], rtol=1e-5, atol=1e-6)
    safe = g["margin"] > 1e-4
    assert safe.mean() > 0.95
    np.testing.assert_array_equal(a[safe], g["actions"][safe])
    lp, g_lp = lp[safe], g["logp"][safe]
    np.testing.assert_allclose(lp, g_lp, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(P.values(w, g["obs"], g["feat"], g["cnt"]), g["values"], rtol=1e-12, atol=1e-12)
    ret, adv = P.gae(g["gae_reward"], g["gae_cost"], g["gae_values"], g["gae_done"], 0.99, 0.95)
    np.testing.assert_allclose(ret, g["gae_returns"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(adv, g["gae_adv"], rtol=1e-12, atol=1e-12)
    # the weights in the fixture are the ones GraphAttentionActor(5, seed=2026) draws (init is part of the pin)
    w2 = P.weights_from_state_dict(GraphAttentionActor(5, seed=2026).state_dict())
    for k in w:
        np.testing.assert_array_equal(w[k], w2[k])
