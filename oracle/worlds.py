"""The oracle's OWN description of a world: data types, buffer shapes and the literal constant
tables of the test worlds.  Nothing here imports `gs_marl_b200` — deleting the product package
does not stop `oracle/` from producing the golden vectors (tests/test_oracle_env.py checks that in
a subprocess with the import blocked), and tests/test_oracle_env.py::
test_oracle_tables_match_product_scenarios compares these tables with what the product's scenario
files build, field by field, so a wrong slot table / shape / preset on EITHER side is a test
failure instead of a shared input.

TEST INFRASTRUCTURE ONLY — PARITY UNPINNED.  The tables restate SPEC.md §1 and the UNVERIFIED
lineage constants of SURVEY.md Appendix A; none of it comes from the withheld GS-MARL sources
(reference readme.md:1).  tools/unblock.py regenerates them from a live `gsmarl` import when the
sources appear.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Optional, Sequence

import numpy as np

# ---- SPEC.md §1, §6: enumerations and row widths (literal, restated from the document) ---------
OBS_DIM = 6            # (v.x, v.y, p.x, p.y, target.x - p.x, target.y - p.y)
NBR_FEAT_DIM = 6       # (d.x, d.y, dv.x, dv.y, dist, type)
MAX_LSA_N = 32
F32, F64 = 0, 1
SCN = {"navigation": 0, "polygon": 1, "line": 2}
ACT = {"discrete": 0, "continuous": 1}
ENT_AGENT, ENT_GOAL, ENT_OBSTACLE, ENT_MARKER = 0, 1, 2, 3
ORC_ABI_VERSION = 7    # layout generation of orc_types.h (kept equal to the library's for convenience)

_dp = C.POINTER(C.c_double)


class OrcConfig(C.Structure):
    """ctypes mirror of `orc_config` (oracle/orc_types.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32),
        ("dtype", C.c_int32), ("scenario", C.c_int32), ("action_mode", C.c_int32),
        ("n_agents", C.c_int32), ("n_landmarks", C.c_int32), ("max_nbrs", C.c_int32),
        ("episode_length", C.c_int32), ("n_discrete_actions", C.c_int32),
        ("share_reward", C.c_int32), ("cost_obstacles", C.c_int32),
        ("own_goal_always", C.c_int32), ("reserved0", C.c_int32),
        ("dt", C.c_double), ("damping", C.c_double), ("contact_force", C.c_double),
        ("contact_margin", C.c_double), ("sensing_radius", C.c_double),
        ("w_dist", C.c_double), ("w_goal", C.c_double), ("goal_tol", C.c_double),
        ("polygon_radius", C.c_double), ("spawn_extent", C.c_double * 4),
        ("discrete_u", _dp), ("size", _dp), ("collide", C.POINTER(C.c_uint8)),
        ("type", C.POINTER(C.c_int32)), ("mass", _dp), ("accel", _dp), ("max_speed", _dp),
        ("slot_table", _dp),
    ]


class OrcStepIO(C.Structure):
    """ctypes mirror of `orc_step_io` (oracle/orc_types.h)."""
    FIELDS = ("actions", "obs", "nbr_idx", "nbr_feat", "nbr_cnt", "adj", "reward", "cost",
              "done", "assign")
    _fields_ = [(k, C.c_void_p) for k in FIELDS]


# Fields a caller chooses (the INPUTS of both implementations).  Everything else — shapes, adjacency
# words, the slot tables — is derived here from SPEC.md, never taken from the product.
INPUT_FIELDS = ("dtype", "scenario", "action_mode", "n_agents", "n_landmarks", "max_nbrs",
                "episode_length", "share_reward", "cost_obstacles", "own_goal_always", "dt", "damping",
                "contact_force", "contact_margin", "sensing_radius", "w_dist", "w_goal", "goal_tol",
                "polygon_radius", "spawn_extent", "discrete_u", "size", "collide", "type", "mass",
                "accel", "max_speed")


def spec_slot_table(scenario: str, n_agents: int):
    """SPEC §1: POLYGON slot k = (cos, sin)(2 pi k / N) in fp64; LINE slot k = fraction (k+1)/(N+1)."""
    if scenario == "polygon":
        return tuple((math.cos(2.0 * math.pi * k / n_agents), math.sin(2.0 * math.pi * k / n_agents))
                     for k in range(n_agents))
    if scenario == "line":
        return tuple(((k + 1.0) / (n_agents + 1.0), 0.0) for k in range(n_agents))
    return None


@dataclasses.dataclass(frozen=True)
class World:
    dtype: str
    scenario: str
    action_mode: str
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
    spawn_extent: Sequence[float]
    discrete_u: Sequence[Sequence[float]]
    size: Sequence[float]
    collide: Sequence[int]
    type: Sequence[int]
    mass: Sequence[float]
    accel: Sequence[float]
    max_speed: Sequence[float]
    slot_table: Optional[Sequence[Sequence[float]]] = None

    @property
    def n_entities(self) -> int:
        return self.n_agents + self.n_landmarks

    @property
    def adj_words(self) -> int:                      # SPEC §6: bit e%32 of word e/32
        return (self.n_entities + 31) // 32

    @property
    def np_real(self):
        return np.float32 if self.dtype == "f32" else np.float64

    def replace(self, **kw) -> "World":
        return dataclasses.replace(self, **kw)

    def io_shapes(self, n_envs: int) -> dict:
        """SPEC §6-7 buffer shapes, restated (not imported from the product)."""
        N, K, r = self.n_agents, self.max_nbrs, self.np_real
        return {
            "actions": ((np.int32, (n_envs, N)) if self.action_mode == "discrete" else (r, (n_envs, N, 2))),
            "obs": (r, (n_envs, N, OBS_DIM)),
            "nbr_idx": (np.int32, (n_envs, N, K)),
            "nbr_feat": (r, (n_envs, N, K, NBR_FEAT_DIM)),
            "nbr_cnt": (np.int32, (n_envs, N)),
            "adj": (np.uint32, (n_envs, N, self.adj_words)),
            "reward": (r, (n_envs, N)),
            "cost": (r, (n_envs, N)),
            "done": (np.uint8, (n_envs, N)),
            "assign": (np.int32, (n_envs, N)),
        }

    def to_c(self):
        """(OrcConfig, keepalive)."""
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
        assert arr["size"].shape == (E,) and arr["collide"].shape == (E,) and arr["type"].shape == (E,)
        assert arr["mass"].shape == (N,) and arr["accel"].shape == (N,) and arr["max_speed"].shape == (N,)
        c = OrcConfig()
        c.struct_size, c.abi_version = C.sizeof(OrcConfig), ORC_ABI_VERSION
        c.dtype = F32 if self.dtype == "f32" else F64
        c.scenario, c.action_mode = SCN[self.scenario], ACT[self.action_mode]
        c.n_agents, c.n_landmarks, c.max_nbrs = self.n_agents, self.n_landmarks, self.max_nbrs
        c.episode_length = self.episode_length
        c.n_discrete_actions = arr["discrete_u"].shape[0]
        c.share_reward, c.cost_obstacles = int(self.share_reward), int(self.cost_obstacles)
        c.own_goal_always = int(self.own_goal_always)
        for k in ("dt", "damping", "contact_force", "contact_margin", "sensing_radius", "w_dist", "w_goal",
                  "goal_tol", "polygon_radius"):
            setattr(c, k, float(getattr(self, k)))
        for i in range(4):
            c.spawn_extent[i] = float(self.spawn_extent[i])
        c.discrete_u = arr["discrete_u"].ctypes.data_as(_dp)
        c.size = arr["size"].ctypes.data_as(_dp)
        c.collide = arr["collide"].ctypes.data_as(C.POINTER(C.c_uint8))
        c.type = arr["type"].ctypes.data_as(C.POINTER(C.c_int32))
        c.mass, c.accel = arr["mass"].ctypes.data_as(_dp), arr["accel"].ctypes.data_as(_dp)
        c.max_speed = arr["max_speed"].ctypes.data_as(_dp)
        if self.slot_table is not None:
            arr["slot_table"] = np.ascontiguousarray(self.slot_table, dtype=np.float64)
            assert arr["slot_table"].shape == (N, 2)
            c.slot_table = arr["slot_table"].ctypes.data_as(_dp)
        return c, arr


def as_world(cfg) -> World:
    """An oracle World from any object carrying the INPUT fields (e.g. the product's WorldConfig,
    which a parity test hands to both sides the way it hands both the same actions).  Derived
    tables are NOT copied: the slot table is recomputed from SPEC §1 and must agree with the
    caller's, shapes come from this module."""
    if isinstance(cfg, World):
        return cfg
    kw = {k: getattr(cfg, k) for k in INPUT_FIELDS}
    for k in ("spawn_extent", "size", "collide", "type", "mass", "accel", "max_speed"):
        kw[k] = tuple(kw[k])
    kw["discrete_u"] = tuple(tuple(u) for u in kw["discrete_u"])
    own = spec_slot_table(kw["scenario"], kw["n_agents"])
    theirs = getattr(cfg, "slot_table", None)
    if (own is None) != (theirs is None) or (
            own is not None and not np.array_equal(np.asarray(own, np.float64), np.asarray(theirs, np.float64))):
        raise AssertionError("slot_table handed to the oracle differs from SPEC.md §1's table for "
                             f"{kw['scenario']} N={kw['n_agents']}")
    return World(slot_table=own, **kw)


# ---- literal constant tables of the test worlds (UNVERIFIED lineage values) --------
_WORLD = dict(dt=0.1, damping=0.25, contact_force=100.0, contact_margin=0.001)
_REWARD = dict(w_dist=1.0, w_goal=1.0, goal_tol=0.1)
_DISCRETE_U = ((0.0, 0.0), (1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0))
_AGENT = dict(size=0.10, mass=1.0, accel=5.0, max_speed=1.3)
_GOAL_SIZE, _OBSTACLE_SIZE, _MARKER_SIZE = 0.05, 0.16, 0.16
_SENSING_RADIUS = 1.0
_EPISODE = {"navigation": 25, "polygon": 100, "line": 100}
_ALIASES = {"simple_formation": "polygon", "simple_line": "line"}


def make_world(name: str, n_agents: int, *, dtype: str, n_obstacles=None, action_mode="discrete",
               max_nbrs=None, episode_length=None, sensing_radius=None, share_reward=False,
               cost_obstacles=None, own_goal_always=None, polygon_radius=0.5, **overrides) -> World:
    """The oracle's literal table for world `name` with N agents (same keyword surface as the
    product's scenarios.load(name).make_world, written independently)."""
    scn = _ALIASES.get(name, name)
    N = int(n_agents)
    ext = math.sqrt(max(N, 3) / 3.0)
    if scn == "navigation":
        n_obs = N if n_obstacles is None else int(n_obstacles)
        L = N + n_obs
        size = [_AGENT["size"]] * N + [_GOAL_SIZE] * N + [_OBSTACLE_SIZE] * n_obs
        collide = [1] * N + [0] * N + [1] * n_obs
        etype = [ENT_AGENT] * N + [ENT_GOAL] * N + [ENT_OBSTACLE] * n_obs
        spawn = (ext, ext, ext, ext)
        cost_obstacles = True if cost_obstacles is None else cost_obstacles
        own_goal_always = True if own_goal_always is None else own_goal_always
        polygon_radius = 0.0
    elif scn == "polygon":
        L = 1
        size = [_AGENT["size"]] * N + [_MARKER_SIZE]
        collide = [1] * N + [0]
        etype = [ENT_AGENT] * N + [ENT_MARKER]
        spawn = (ext, ext, ext, 0.5 * ext)
        cost_obstacles, own_goal_always = False, False
    elif scn == "line":
        L = 2
        size = [_AGENT["size"]] * N + [_MARKER_SIZE] * 2
        collide = [1] * N + [0, 0]
        etype = [ENT_AGENT] * N + [ENT_MARKER] * 2
        spawn = (ext, ext, ext, ext)
        cost_obstacles, own_goal_always = False, False
        polygon_radius = 0.0
    else:
        raise KeyError(f"oracle has no table for world {name!r}")
    E = N + L
    if max_nbrs is None:
        max_nbrs = E - 1 if E - 1 <= 8 else min(32, (E - 1) // 4 * 4)
    kw = dict(dtype=dtype, scenario=scn, action_mode=action_mode, n_agents=N, n_landmarks=L,
              max_nbrs=max_nbrs, episode_length=_EPISODE[scn] if episode_length is None else episode_length,
              share_reward=share_reward, cost_obstacles=cost_obstacles, own_goal_always=own_goal_always,
              sensing_radius=_SENSING_RADIUS if sensing_radius is None else sensing_radius,
              polygon_radius=polygon_radius, spawn_extent=spawn, discrete_u=_DISCRETE_U,
              size=tuple(size), collide=tuple(collide), type=tuple(etype),
              mass=(_AGENT["mass"],) * N, accel=(_AGENT["accel"],) * N, max_speed=(_AGENT["max_speed"],) * N,
              slot_table=spec_slot_table(scn, N), **_WORLD, **_REWARD)
    kw.update(overrides)
    return World(**kw)
