"""Vector-env boundary — counterpart of gsmarl/envs/mpe_env/env_wrappers.py
(GSMARL.egg-info/SOURCES.txt:11), which (by lineage) steps `n_rollout_threads` numpy envs
in subprocess workers and stacks their results along axis 0.

`GraphVecEnv` is the numpy-facing drop-in: HOST arrays in, HOST arrays out, the host<->
device copies inside the call (C ABI `gsm_step_host` over one pinned arena).  No torch.
`shard_bounds` / `ShardedStats` are the multi-GPU plumbing: env instances are independent,
so ranks own disjoint contiguous env ranges and nothing is exchanged on the step path;
only the final counters are summed (SURVEY.md §8 e).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import abi
from .config import WorldConfig


def shard_bounds(n_envs_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) env range of `rank`; the first n % world ranks get one more."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    q, r = divmod(int(n_envs_total), world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class GraphVecEnv:
    """numpy drop-in over gsm_*_host.  Returned arrays are views of the pinned arena and
    are overwritten by the next call (copy them to keep them).  The output views are READ-ONLY: the
    sparse export (`set_host_outputs`) sends only what differs from what the arena already holds."""

    def __init__(self, world: WorldConfig, n_envs: int, device: int = 0, env_offset: int = 0,
                 auto_reset: bool = False, seed: int = 0):
        self.world, self.n_envs, self.auto_reset = world, int(n_envs), bool(auto_reset)
        self.lib = abi.load_library()
        self._c, self._keep = world.to_c()
        self._h = C.c_void_p()
        abi.check(self.lib, self.lib.gsm_create(C.byref(self._c), self.n_envs, int(env_offset),
                                                int(device), C.byref(self._h)))
        self._seed = int(seed)
        self._io = abi.GsmStepIO()
        abi.check(self.lib, self.lib.gsm_host_io(self._h, C.byref(self._io)), self._h)
        self.buf = {}
        for k, (dt, shape) in world.io_shapes(self.n_envs).items():
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            raw = (C.c_char * nbytes).from_address(getattr(self._io, k))
            self.buf[k] = np.frombuffer(raw, dtype=dt).reshape(shape)
            if k != "actions":
                self.buf[k].flags.writeable = False
        self.num_envs = self.n_envs
        self.n = world.n_agents

    def _check(self, st):
        abi.check(self.lib, st, self._h)

    _IO_ORDER = ("actions", "obs", "nbr_idx", "nbr_feat", "nbr_cnt", "adj", "reward", "cost", "done", "assign")

    def set_host_outputs(self, outputs=None, sparse: bool = True):
        """Which outputs the *_host calls deliver (None: all) and how (`gsm_set_host_outputs`):
        sparse=True sends only the nbr_cnt valid rows of nbr_feat over PCIe (the host arrays
        stay bit-identical to the device tensors), sparse=False is one dense D2H copy.  Outputs left out
        keep whatever the host array held."""
        mask = 0xFFFFFFFF
        if outputs is not None:
            unknown = set(outputs) - set(self._IO_ORDER)
            if unknown:
                raise ValueError(f"unknown outputs {sorted(unknown)}")
            mask = sum(1 << self._IO_ORDER.index(k) for k in outputs) | 1
        self._check(self.lib.gsm_set_host_outputs(self._h, mask, int(bool(sparse))))

    def seed(self, seed: int):
        self._seed = int(seed)

    def _graph(self):
        b = self.buf
        return {"nbr_idx": b["nbr_idx"], "nbr_feat": b["nbr_feat"], "nbr_cnt": b["nbr_cnt"],
                "adj": b["adj"]}

    def reset(self, mask: Optional[np.ndarray] = None):
        m = None
        if mask is not None:
            self._mask = np.ascontiguousarray(mask, dtype=np.uint8)
            m = self._mask.ctypes.data
        self._check(self.lib.gsm_reset_host(self._h, self._seed, m, 1, C.byref(self._io)))
        return self.buf["obs"], self._graph()

    def step(self, actions):
        np.copyto(self.buf["actions"], np.asarray(actions).reshape(self.buf["actions"].shape),
                  casting="same_kind")
        self._check(self.lib.gsm_step_host(self._h, C.byref(self._io)))
        b = self.buf
        if self.auto_reset and b["done"].any():
            # the re-draw refreshes obs / graph of the finished envs; reward, cost and done are not outputs of a
            # reset and keep the terminal step's values in the arena
            done = b["done"].copy()
            self._check(self.lib.gsm_reset_host(self._h, self._seed, done.ctypes.data,
                                                self.world.n_agents, C.byref(self._io)))
        infos = {"assign": b["assign"], "collisions": b["cost"]}
        return b["obs"], self._graph(), b["reward"], b["cost"], b["done"], infos

    def set_state(self, agent_state=None, landmark_pos=None, step_count=None):
        r = self.world.np_real
        a = None if agent_state is None else np.ascontiguousarray(agent_state, r)
        l = None if landmark_pos is None else np.ascontiguousarray(landmark_pos, r)
        t = None if step_count is None else np.ascontiguousarray(step_count, np.int32)
        ptr = [x.ctypes.data if x is not None and x.size else None for x in (a, l, t)]
        self._check(self.lib.gsm_set_state_host(self._h, *ptr))

    def get_state(self):
        w, r = self.world, self.world.np_real
        a = np.zeros((self.n_envs, w.n_agents, 4), r)
        l = np.zeros((self.n_envs, w.n_landmarks, 2), r)
        t = np.zeros((self.n_envs,), np.int32)
        self._check(self.lib.gsm_get_state_host(self._h, a.ctypes.data,
                                                l.ctypes.data if l.size else None, t.ctypes.data))
        return a, l, t

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.gsm_kernel_launches(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.buf = {}
            self.lib.gsm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedStats:
    """Final stats gather of an env-sharded rollout: the only collective of the path.
    Works with any initialised torch.distributed backend (nccl on GPUs, gloo on CPU)."""

    FIELDS = ("env_steps", "agent_steps", "reward_sum", "cost_sum", "done_count")

    def __init__(self):
        self.v = {k: 0.0 for k in self.FIELDS}

    def add(self, n_envs: int, n_agents: int, reward_sum: float, cost_sum: float, done_count: float):
        self.v["env_steps"] += n_envs
        self.v["agent_steps"] += n_envs * n_agents
        self.v["reward_sum"] += float(reward_sum)
        self.v["cost_sum"] += float(cost_sum)
        self.v["done_count"] += float(done_count)

    def all_reduce(self, device=None) -> dict:
        import torch
        import torch.distributed as dist
        t = torch.tensor([self.v[k] for k in self.FIELDS], dtype=torch.float64, device=device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return dict(zip(self.FIELDS, t.tolist()))
