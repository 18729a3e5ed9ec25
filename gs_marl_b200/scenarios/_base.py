"""BaseScenario — counterpart of multiagent/scenario.py (SOURCES.txt:19)."""
from __future__ import annotations

from .. import abi
from ..config import WorldConfig
from .. import presets as P


class BaseScenario:
    name = "base"

    def make_world(self, **kw) -> WorldConfig:  # pragma: no cover - interface
        raise NotImplementedError

    # The reference's per-agent callbacks are fused into the step kernel; these names
    # exist so a reader of the reference finds where each one went.
    reset_world = "SPEC.md §8 -> gsm_reset"
    observation = "SPEC.md §6 -> gsm_step / gsm_observe"
    reward = "SPEC.md §7 -> gsm_step"
    cost = "SPEC.md §7 -> gsm_step"

    @staticmethod
    def _common(n_agents, n_landmarks, dtype, action_mode, max_nbrs, episode_length,
                sensing_radius, share_reward, cost_obstacles, own_goal_always):
        E = n_agents + n_landmarks
        if max_nbrs is None:
            # [DECL] default padded row count: every other entity for small teams, else capped at
            # 32 and kept a multiple of 4 (16-byte TMA bulk-store granularity of the row blocks)
            max_nbrs = E - 1 if E - 1 <= 8 else min(32, (E - 1) // 4 * 4)
        return dict(
            dtype=dtype, action_mode=action_mode, n_agents=n_agents, n_landmarks=n_landmarks,
            max_nbrs=max_nbrs, episode_length=episode_length,
            share_reward=share_reward, cost_obstacles=cost_obstacles,
            own_goal_always=own_goal_always,
            sensing_radius=(P.unverified_sensing_radius(n_agents) if sensing_radius is None
                            else sensing_radius),
            discrete_u=P.UNVERIFIED_DISCRETE_U,
            mass=[P.UNVERIFIED_AGENT["mass"]] * n_agents,
            accel=[P.UNVERIFIED_AGENT["accel"]] * n_agents,
            max_speed=[P.UNVERIFIED_AGENT["max_speed"]] * n_agents,
            **P.UNVERIFIED_WORLD, **P.UNVERIFIED_REWARD,
        )


AGENT, GOAL, OBSTACLE, MARKER = (abi.GSM_ENT_AGENT, abi.GSM_ENT_GOAL, abi.GSM_ENT_OBSTACLE,
                                 abi.GSM_ENT_MARKER)
