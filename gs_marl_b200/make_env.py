"""Counterpart of gsmarl/envs/mpe_env/make_env.py (GSMARL.egg-info/SOURCES.txt:12), the factory
`train_mpe.py` (SOURCES.txt:31) calls to get the vectorised env the runner holds.

The reference file is withheld, so the argument names follow the on-policy lineage's `all_args`
namespace (`scenario_name`, `num_agents`, `n_rollout_threads`, `episode_length`, `seed`) and are
[DECL]: adapt the attribute names when the real `config.py` (SOURCES.txt:8) is visible.  Where the
reference spawns `n_rollout_threads` subprocess workers with one Python env each, this returns ONE
object holding `n_rollout_threads` worlds on a GPU (or this rank's shard of them).

    envs = make_train_env(all_args)                    # numpy in / numpy out (GraphVecEnv)
    envs = make_train_env(all_args, backend="torch")   # CUDA tensors in / out, no PCIe on the step path
"""
from __future__ import annotations

from typing import Any

from . import scenarios
from .env_wrappers import shard_bounds

_BACKENDS = ("numpy", "torch")


def world_from_args(all_args: Any):
    """`scenario.make_world(args)` of the reference: scenario + team size + episode length -> WorldConfig."""
    name = getattr(all_args, "scenario_name", None)
    if name is None:
        raise ValueError("all_args.scenario_name is required")
    n_agents = int(getattr(all_args, "num_agents", 0))
    if n_agents < 1:
        raise ValueError("all_args.num_agents must be >= 1")
    kw = {}
    if getattr(all_args, "episode_length", None) is not None:
        kw["episode_length"] = int(all_args.episode_length)
    if getattr(all_args, "max_nbrs", None) is not None:
        kw["max_nbrs"] = int(all_args.max_nbrs)
    dtype = "f64" if getattr(all_args, "verification_mode", False) else "f32"
    return scenarios.load(name).make_world(n_agents, dtype=dtype, **kw)


def make_train_env(all_args: Any, backend: str = "numpy", device: int = 0, rank: int = 0, world_size: int = 1):
    """The env object for `n_rollout_threads` worlds; with world_size > 1, this rank's contiguous
    shard (reset draws are keyed by the GLOBAL env index, so the union equals one big env)."""
    if backend not in _BACKENDS:
        raise ValueError(f"backend must be one of {_BACKENDS}")
    n_total = int(getattr(all_args, "n_rollout_threads", 0))
    if n_total < 1:
        raise ValueError("all_args.n_rollout_threads must be >= 1")
    if not (0 <= rank < world_size):
        raise ValueError("need 0 <= rank < world_size")
    world = world_from_args(all_args)
    lo, hi = shard_bounds(n_total, world_size, rank)
    seed = int(getattr(all_args, "seed", 0))
    if backend == "numpy":
        from .env_wrappers import GraphVecEnv
        return GraphVecEnv(world, hi - lo, device=device, env_offset=lo, auto_reset=True, seed=seed)
    from .environment import MultiAgentGraphConstrainEnv
    return MultiAgentGraphConstrainEnv(world, hi - lo, device=device, env_offset=lo, auto_reset=True, seed=seed)


def make_eval_env(all_args: Any, backend: str = "numpy", device: int = 0):
    """`n_eval_rollout_threads` worlds with a different seed offset (lineage convention)."""
    import copy
    a = copy.copy(all_args)
    a.n_rollout_threads = int(getattr(all_args, "n_eval_rollout_threads", 1))
    a.seed = int(getattr(all_args, "seed", 0)) * 50000 + 1
    return make_train_env(a, backend=backend, device=device)
