"""Host-side logic: scenario builders, byte accounting, sharding, and the world_size-2
gloo run of the N>1 path (env shards are independent; only final stats are reduced)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gs_marl_b200 import abi, scenarios
from gs_marl_b200.env_wrappers import ShardedStats, shard_bounds
from tests._util import make_cfg, random_actions


def test_scenarios_registry_and_shapes():
    nav = scenarios.load("navigation").make_world(3, dtype="f32")
    assert (nav.n_agents, nav.n_landmarks, nav.n_entities, nav.max_nbrs) == (3, 6, 9, 8)
    assert list(nav.type) == [0] * 3 + [1] * 3 + [2] * 3
    poly = scenarios.load("simple_formation").make_world(6, dtype="f64")
    assert poly.n_landmarks == 1 and poly.polygon_radius == 0.5 and len(poly.slot_table) == 6
    line = scenarios.load("simple_line").make_world(4, dtype="f64")
    assert line.n_landmarks == 2 and line.slot_table[0][0] == pytest.approx(0.2)
    with pytest.raises(KeyError):
        scenarios.load("nope")
    sh = nav.io_shapes(10)
    assert sh["nbr_feat"][1] == (10, 3, 8, abi.GSM_NBR_FEAT_DIM) and sh["adj"][1] == (10, 3, 1)
    big = scenarios.load("navigation").make_world(96, dtype="f32", max_nbrs=32)
    assert big.adj_words == 9


def test_world_config_has_no_defaults():
    import dataclasses
    from gs_marl_b200.config import WorldConfig
    for f in dataclasses.fields(WorldConfig):
        assert f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING, f.name


def test_bytes_per_agent_step():
    nav = make_cfg("navigation", 3, "f32")
    # state r/w 32 + action 4 + obs 24 + idx 32 + feat 192 + cnt 4 + adj 4 + rew 4 + cost 4
    # + done 1 + assign 4 = 305, plus per-env (6 landmarks*8 + 8)/3 -> 19
    assert nav.bytes_per_agent_step() == 305 + 19


def test_shard_bounds_partition():
    for n, w in [(16384, 8), (10, 4), (7, 8), (65536, 8)]:
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gsm_oracle as O          # tests may use the oracle as the env backend
    cfg = make_cfg("navigation", 3, "f64")
    lo, hi = shard_bounds(n_total, world, rank)
    env = O.OracleEnv(cfg, hi - lo, env_offset=lo)
    env.reset(11)
    rng = np.random.default_rng(0)
    acts = random_actions(cfg, rng, (5, n_total))[:, lo:hi]
    stats = ShardedStats()
    for t in range(5):
        o = env.step(acts[t])
        stats.add(hi - lo, cfg.n_agents, o["reward"].sum(), o["cost"].sum(), o["done"].sum())
    tot = stats.all_reduce()
    gathered = [None] * world
    dist.all_gather_object(gathered, env.agent_state)
    if rank == 0:
        q.put((tot, np.concatenate(gathered)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_rollout_equals_single():
    from oracle import gsm_oracle as O
    n_total, world = 10, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    tot, state = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    cfg = make_cfg("navigation", 3, "f64")
    env = O.OracleEnv(cfg, n_total)
    env.reset(11)
    acts = random_actions(cfg, np.random.default_rng(0), (5, n_total))
    rs = cs = 0.0
    for t in range(5):
        o = env.step(acts[t])
        rs += o["reward"].sum()
        cs += o["cost"].sum()
    assert (state == env.agent_state).all()                 # sharding is invisible, bit-exact
    assert tot["env_steps"] == 5 * n_total and tot["agent_steps"] == 5 * n_total * 3
    assert tot["reward_sum"] == pytest.approx(rs, rel=1e-12) and tot["cost_sum"] == cs


def test_make_env_argument_mapping_and_no_fallback():
    """make_env.py mirror: all_args -> WorldConfig mapping is host logic; creating the env without a
    device raises (never a CPU env)."""
    import types

    import torch

    from gs_marl_b200 import abi, make_env
    args = types.SimpleNamespace(scenario_name="polygon", num_agents=6, n_rollout_threads=10, episode_length=100, seed=3)
    w = make_env.world_from_args(args)
    assert w.n_agents == 6 and w.episode_length == 100 and w.dtype == "f32"
    assert make_env.world_from_args(types.SimpleNamespace(scenario_name="navigation", num_agents=3,
                                                          verification_mode=True)).dtype == "f64"
    with pytest.raises(KeyError):
        make_env.world_from_args(types.SimpleNamespace(scenario_name="simple_spread", num_agents=3))
    with pytest.raises(ValueError):
        make_env.world_from_args(types.SimpleNamespace(scenario_name="navigation", num_agents=0))
    with pytest.raises(ValueError):
        make_env.make_train_env(args, backend="jax")
    with pytest.raises(ValueError):
        make_env.make_train_env(types.SimpleNamespace(scenario_name="navigation", num_agents=3, n_rollout_threads=0))
    if not torch.cuda.is_available():
        for backend in ("numpy", "torch"):
            with pytest.raises(abi.GsmError):
                make_env.make_train_env(args, backend=backend)
