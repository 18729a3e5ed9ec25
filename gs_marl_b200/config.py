"""WorldConfig — the explicit, default-free description of a world + scenario.

Stands in for what `scenario.make_world(args)` builds in the reference
(gsmarl/envs/mpe_env/multiagent/scenarios/*.py, GSMARL.egg-info/SOURCES.txt:21-25) plus
the `World` constants of core.py (SOURCES.txt:14).  Those sources are withheld, so no
field has a default: a caller must state every constant.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional, Sequence

import numpy as np

from . import abi


@dataclasses.dataclass(frozen=True)
class WorldConfig:
    dtype: str                      # "f32" (production) | "f64" (verification)
    scenario: str                   # "navigation" | "polygon" | "line"
    action_mode: str                # "discrete" | "continuous"
    n_agents: int
    n_landmarks: int
    max_nbrs: int
    episode_length: int
    share_reward: bool
    cost_obstacles: bool
    own_goal_always: bool
    dt: float
    damping: float
    contact_force: float
    contact_margin: float
    sensing_radius: float
    w_dist: float
    w_goal: float
    goal_tol: float
    polygon_radius: float
    spawn_extent: Sequence[float]   # per entity type (agent, goal, obstacle, marker)
    discrete_u: Sequence[Sequence[float]]   # [A][2]
    size: Sequence[float]           # [N+L]
    collide: Sequence[int]          # [N+L]
    type: Sequence[int]             # [N+L]
    mass: Sequence[float]           # [N]
    accel: Sequence[float]          # [N]
    max_speed: Sequence[float]      # [N]
    slot_table: Optional[Sequence[Sequence[float]]]  # [N][2] or None (navigation)

    _DT = {"f32": abi.GSM_F32, "f64": abi.GSM_F64}
    _SC = {"navigation": abi.GSM_SCN_NAVIGATION, "polygon": abi.GSM_SCN_POLYGON,
           "line": abi.GSM_SCN_LINE}
    _AM = {"discrete": abi.GSM_ACT_DISCRETE, "continuous": abi.GSM_ACT_CONTINUOUS}

    @property
    def n_entities(self) -> int:
        return self.n_agents + self.n_landmarks

    @property
    def adj_words(self) -> int:
        return (self.n_entities + 31) // 32

    @property
    def np_real(self):
        return np.float32 if self.dtype == "f32" else np.float64

    def replace(self, **kw) -> "WorldConfig":
        return dataclasses.replace(self, **kw)

    def to_c(self):
        """Returns (GsmConfig, keepalive) — keepalive owns the host arrays."""
        N, E = self.n_agents, self.n_entities
        arr = {
            "discrete_u": np.ascontiguousarray(self.discrete_u, dtype=np.float64).reshape(-1, 2),
            "size": np.ascontiguousarray(self.size, dtype=np.float64),
            "collide": np.ascontiguousarray(self.collide, dtype=np.uint8),
            "type": np.ascontiguousarray(self.type, dtype=np.int32),
            "mass": np.ascontiguousarray(self.mass, dtype=np.float64),
            "accel": np.ascontiguousarray(self.accel, dtype=np.float64),
            "max_speed": np.ascontiguousarray(self.max_speed, dtype=np.float64),
        }
        for k, n in (("size", E), ("collide", E), ("type", E), ("mass", N), ("accel", N),
                     ("max_speed", N)):
            if arr[k].shape != (n,):
                raise ValueError(f"{k} must have shape ({n},), got {arr[k].shape}")
        if self.slot_table is not None:
            arr["slot_table"] = np.ascontiguousarray(self.slot_table, dtype=np.float64)
            if arr["slot_table"].shape != (N, 2):
                raise ValueError("slot_table must be [n_agents][2]")
        c = abi.GsmConfig()
        c.struct_size = C.sizeof(abi.GsmConfig)
        c.abi_version = abi.GSM_ABI_VERSION
        c.dtype = self._DT[self.dtype]
        c.scenario = self._SC[self.scenario]
        c.action_mode = self._AM[self.action_mode]
        c.n_agents, c.n_landmarks, c.max_nbrs = self.n_agents, self.n_landmarks, self.max_nbrs
        c.episode_length = self.episode_length
        c.n_discrete_actions = arr["discrete_u"].shape[0]
        c.share_reward = int(self.share_reward)
        c.cost_obstacles = int(self.cost_obstacles)
        c.own_goal_always = int(self.own_goal_always)
        for k in ("dt", "damping", "contact_force", "contact_margin", "sensing_radius", "w_dist",
                  "w_goal", "goal_tol", "polygon_radius"):
            setattr(c, k, float(getattr(self, k)))
        if len(self.spawn_extent) != 4:
            raise ValueError("spawn_extent needs 4 entries (agent, goal, obstacle, marker)")
        for i in range(4):
            c.spawn_extent[i] = float(self.spawn_extent[i])
        c.discrete_u = arr["discrete_u"].ctypes.data_as(C.POINTER(C.c_double))
        c.size = arr["size"].ctypes.data_as(C.POINTER(C.c_double))
        c.collide = arr["collide"].ctypes.data_as(C.POINTER(C.c_uint8))
        c.type = arr["type"].ctypes.data_as(C.POINTER(C.c_int32))
        c.mass = arr["mass"].ctypes.data_as(C.POINTER(C.c_double))
        c.accel = arr["accel"].ctypes.data_as(C.POINTER(C.c_double))
        c.max_speed = arr["max_speed"].ctypes.data_as(C.POINTER(C.c_double))
        if self.slot_table is not None:
            c.slot_table = arr["slot_table"].ctypes.data_as(C.POINTER(C.c_double))
        return c, arr

    # ---- shapes of the per-step buffers (SPEC.md §6-7) -------------------------------
    def io_shapes(self, n_envs: int) -> dict:
        N, K = self.n_agents, self.max_nbrs
        r = self.np_real
        act = (np.int32, (n_envs, N)) if self.action_mode == "discrete" else (r, (n_envs, N, 2))
        return {
            "actions": act,
            "obs": (r, (n_envs, N, abi.GSM_OBS_DIM)),
            "nbr_idx": (np.int32, (n_envs, N, K)),
            "nbr_feat": (r, (n_envs, N, K, abi.GSM_NBR_FEAT_DIM)),
            "nbr_cnt": (np.int32, (n_envs, N)),
            "adj": (np.uint32, (n_envs, N, self.adj_words)),
            "reward": (r, (n_envs, N)),
            "cost": (r, (n_envs, N)),
            "done": (np.uint8, (n_envs, N)),
            "assign": (np.int32, (n_envs, N)),
        }

    def bytes_per_agent_step(self) -> int:
        """Algorithmic HBM bytes per agent-step of gsm_step: state read +
        write, action read, every output written once; landmark reads amortised per env."""
        rb = 4 if self.dtype == "f32" else 8
        N, L, K = self.n_agents, self.n_landmarks, self.max_nbrs
        act = 4 if self.action_mode == "discrete" else 2 * rb
        per_agent = (2 * 4 * rb + act + abi.GSM_OBS_DIM * rb + K * 4
                     + K * abi.GSM_NBR_FEAT_DIM * rb + 4 + 4 * self.adj_words + rb + rb + 1 + 4)
        per_env = L * 2 * rb + 2 * 4          # landmark positions read, step counter r/w
        return per_agent + (per_env + N - 1) // N

    def bytes_fused(self, n_envs: int, n_steps: int) -> int:
        """Algorithmic HBM bytes of ONE fused `n_steps`-step launch over `n_envs` envs
        (gsm_rollout): the agent state, the landmark positions and the step counter cross HBM
        once per LAUNCH (they live in registers / shared memory in between), the action and every
        output once per STEP.  This is the roofline numerator of bench.py."""
        rb = 4 if self.dtype == "f32" else 8
        N, L, K = self.n_agents, self.n_landmarks, self.max_nbrs
        act = 4 if self.action_mode == "discrete" else 2 * rb
        stream = (act + abi.GSM_OBS_DIM * rb + K * 4 + K * abi.GSM_NBR_FEAT_DIM * rb + 4
                  + 4 * self.adj_words + rb + rb + 1 + 4)           # per agent per step
        once = N * 2 * 4 * rb + L * 2 * rb + 2 * 4                   # per env per launch
        return n_envs * (n_steps * N * stream + once)
