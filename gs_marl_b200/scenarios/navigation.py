"""Cooperative navigation (reference readme.md:44,71-77; demo/navigation/*.gif: N agents,
N goals, N obstacles).  Which reference file holds it (exp1.py / exp2.py,
SOURCES.txt:21-22) is unknown.  Constants: presets.UNVERIFIED_* unless overridden."""
from __future__ import annotations

from ._base import BaseScenario, AGENT, GOAL, OBSTACLE
from .. import presets as P
from ..config import WorldConfig


class Scenario(BaseScenario):
    name = "navigation"

    def make_world(self, n_agents: int, *, dtype: str, n_obstacles=None, action_mode="discrete",
                   max_nbrs=None, episode_length=25, sensing_radius=None, share_reward=False,
                   cost_obstacles=True, own_goal_always=True, **overrides) -> WorldConfig:
        n_obs = n_agents if n_obstacles is None else n_obstacles
        L = n_agents + n_obs
        ext = P.unverified_spawn_extent(n_agents)
        kw = self._common(n_agents, L, dtype, action_mode, max_nbrs, episode_length,
                          sensing_radius, share_reward, cost_obstacles, own_goal_always)
        kw.update(
            scenario="navigation", polygon_radius=0.0, slot_table=None,
            spawn_extent=(ext, ext, ext, ext),
            size=[P.UNVERIFIED_AGENT["size"]] * n_agents + [P.UNVERIFIED_GOAL_SIZE] * n_agents
                 + [P.UNVERIFIED_OBSTACLE_SIZE] * n_obs,
            collide=[1] * n_agents + [0] * n_agents + [1] * n_obs,
            type=[AGENT] * n_agents + [GOAL] * n_agents + [OBSTACLE] * n_obs,
        )
        kw.update(overrides)
        return WorldConfig(**kw)
