"""ctypes front-end of oracle/libgsm_oracle.so (C restatement of SPEC.md).

TEST INFRASTRUCTURE ONLY — PARITY UNPINNED (reference sources withheld, readme.md:1).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import worlds

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libgsm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_DIR, f) for f in ("gsm_oracle.c", "gsm_oracle_impl.h", "orc_types.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s)):
        subprocess.run(["make", "-C", _DIR, "-s", "-B"], check=True, capture_output=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        cfgp, iop = C.POINTER(worlds.OrcConfig), C.POINTER(worlds.OrcStepIO)
        for sfx in ("f64", "f32"):
            getattr(_lib, f"orc_step_{sfx}").argtypes = [cfgp, C.c_int64, C.c_void_p, C.c_void_p,
                                                         C.c_void_p, iop]
            getattr(_lib, f"orc_observe_{sfx}").argtypes = [cfgp, C.c_int64, C.c_void_p,
                                                            C.c_void_p, C.c_void_p, iop]
            getattr(_lib, f"orc_evaluate_{sfx}").argtypes = [cfgp, C.c_int64, C.c_void_p,
                                                             C.c_void_p, C.c_void_p, iop]
            getattr(_lib, f"orc_reset_{sfx}").argtypes = [cfgp, C.c_int64, C.c_int64, C.c_uint64,
                                                          C.c_void_p, C.c_int64, C.c_void_p,
                                                          C.c_void_p, C.c_void_p, C.c_void_p]
            getattr(_lib, f"orc_lsa_{sfx}").argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_set_threads.argtypes = [C.c_int]
        _lib.orc_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
    return _lib


def set_threads(n: int) -> int:
    return lib().orc_set_threads(n)


def philox(c0, c1, c2, c3, k0, k1) -> np.ndarray:
    out = np.zeros(4, np.uint32)
    lib().orc_philox(c0, c1, c2, c3, k0, k1, out.ctypes.data)
    return out


def lsa(cost: np.ndarray) -> np.ndarray:
    """col4row of the n x n problem (batched over leading dims)."""
    cost = np.ascontiguousarray(cost)
    sfx = "f32" if cost.dtype == np.float32 else "f64"
    if sfx == "f64":
        cost = cost.astype(np.float64, copy=False)
    n = cost.shape[-1]
    flat = cost.reshape(-1, n, n)
    out = np.empty((flat.shape[0], n), np.int32)
    fn = getattr(lib(), f"orc_lsa_{sfx}")
    for b in range(flat.shape[0]):
        rc = fn(flat[b].ctypes.data, n, out[b].ctypes.data)
        if rc != 0:
            raise RuntimeError("infeasible")
    return out.reshape(cost.shape[:-1])


class OracleEnv:
    """Batched oracle env holding numpy state; same buffers as gsm_step_io."""

    def __init__(self, cfg, n_envs: int, env_offset: int = 0):
        # `cfg`: an oracle World, or any object carrying the INPUT fields (worlds.as_world re-derives
        # shapes and slot tables from SPEC.md itself and rejects a caller's table that disagrees)
        cfg = worlds.as_world(cfg)
        self.cfg, self.n_envs, self.env_offset = cfg, n_envs, env_offset
        self._c, self._keep = cfg.to_c()
        self._sfx = cfg.dtype
        r = cfg.np_real
        self.agent_state = np.zeros((n_envs, cfg.n_agents, 4), r)
        self.landmark_pos = np.zeros((n_envs, cfg.n_landmarks, 2), r)
        self.step_count = np.zeros(n_envs, np.int32)
        self.episode = np.zeros(n_envs, np.int32)

    def alloc_io(self) -> dict:
        return {k: np.zeros(s, d) for k, (d, s) in self.cfg.io_shapes(self.n_envs).items()}

    @staticmethod
    def _io_struct(bufs: dict) -> worlds.OrcStepIO:
        io = worlds.OrcStepIO()
        for k in worlds.OrcStepIO.FIELDS:
            a = bufs.get(k)
            setattr(io, k, None if a is None else a.ctypes.data)
        return io

    def set_state(self, agent_state=None, landmark_pos=None, step_count=None):
        if agent_state is not None:
            self.agent_state[...] = agent_state
        if landmark_pos is not None:
            self.landmark_pos[...] = landmark_pos
        if step_count is not None:
            self.step_count[...] = step_count

    def reset(self, seed: int, mask=None, mask_stride: int = 1):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        getattr(lib(), f"orc_reset_{self._sfx}")(
            C.byref(self._c), self.n_envs, self.env_offset, seed,
            None if m is None else m.ctypes.data, mask_stride, self.agent_state.ctypes.data,
            self.landmark_pos.ctypes.data, self.step_count.ctypes.data, self.episode.ctypes.data)

    def observe(self, bufs: dict | None = None) -> dict:
        bufs = self.alloc_io() if bufs is None else bufs
        io = self._io_struct(bufs)
        getattr(lib(), f"orc_observe_{self._sfx}")(
            C.byref(self._c), self.n_envs, self.agent_state.ctypes.data,
            self.landmark_pos.ctypes.data, self.step_count.ctypes.data, C.byref(io))
        return bufs

    def evaluate(self, bufs: dict | None = None) -> dict:
        """SPEC §5-7 of the current state INCLUDING reward / cost / done (observe() leaves those)."""
        bufs = self.alloc_io() if bufs is None else bufs
        io = self._io_struct(bufs)
        getattr(lib(), f"orc_evaluate_{self._sfx}")(
            C.byref(self._c), self.n_envs, self.agent_state.ctypes.data,
            self.landmark_pos.ctypes.data, self.step_count.ctypes.data, C.byref(io))
        return bufs

    def step(self, actions: np.ndarray, bufs: dict | None = None) -> dict:
        bufs = self.alloc_io() if bufs is None else bufs
        d, s = self.cfg.io_shapes(self.n_envs)["actions"]
        bufs["actions"] = np.ascontiguousarray(actions, d).reshape(s)
        io = self._io_struct(bufs)
        getattr(lib(), f"orc_step_{self._sfx}")(
            C.byref(self._c), self.n_envs, self.agent_state.ctypes.data,
            self.landmark_pos.ctypes.data, self.step_count.ctypes.data, C.byref(io))
        return bufs
