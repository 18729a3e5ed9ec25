"""Polygon task — scenarios/simple_formation.py in the reference (SOURCES.txt:24;
readme.md:89): N agents around one landmark at the centre of a regular N-gon of radius
0.5, slots assigned by a linear assignment solved every step, cost for agent-agent
collisions."""
from __future__ import annotations

import math

from ._base import BaseScenario, AGENT, MARKER
from .. import presets as P
from ..config import WorldConfig


class Scenario(BaseScenario):
    name = "simple_formation"

    def make_world(self, n_agents: int, *, dtype: str, action_mode="discrete", max_nbrs=None,
                   episode_length=100, sensing_radius=None, share_reward=False,
                   polygon_radius=0.5, **overrides) -> WorldConfig:
        ext = P.unverified_spawn_extent(n_agents)
        kw = self._common(n_agents, 1, dtype, action_mode, max_nbrs, episode_length,
                          sensing_radius, share_reward, False, False)
        kw.update(
            scenario="polygon", polygon_radius=polygon_radius,   # 0.5: readme.md:89
            slot_table=[(math.cos(2.0 * math.pi * k / n_agents),
                         math.sin(2.0 * math.pi * k / n_agents)) for k in range(n_agents)],
            spawn_extent=(ext, ext, ext, 0.5 * ext),
            size=[P.UNVERIFIED_AGENT["size"]] * n_agents + [P.UNVERIFIED_MARKER_SIZE],
            collide=[1] * n_agents + [0],
            type=[AGENT] * n_agents + [MARKER],
        )
        kw.update(overrides)
        return WorldConfig(**kw)
