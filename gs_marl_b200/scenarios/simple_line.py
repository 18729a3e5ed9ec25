"""Line task — scenarios/simple_line.py in the reference (SOURCES.txt:25; readme.md:90):
N agents spread evenly on the segment between two landmarks, slots assigned by a linear
sum assignment solved every step."""
from __future__ import annotations

from ._base import BaseScenario, AGENT, MARKER
from .. import presets as P
from ..config import WorldConfig


class Scenario(BaseScenario):
    name = "simple_line"

    def make_world(self, n_agents: int, *, dtype: str, action_mode="discrete", max_nbrs=None,
                   episode_length=100, sensing_radius=None, share_reward=False,
                   **overrides) -> WorldConfig:
        ext = P.unverified_spawn_extent(n_agents)
        kw = self._common(n_agents, 2, dtype, action_mode, max_nbrs, episode_length,
                          sensing_radius, share_reward, False, False)
        kw.update(
            scenario="line", polygon_radius=0.0,
            slot_table=[((k + 1.0) / (n_agents + 1.0), 0.0) for k in range(n_agents)],
            spawn_extent=(ext, ext, ext, ext),
            size=[P.UNVERIFIED_AGENT["size"]] * n_agents + [P.UNVERIFIED_MARKER_SIZE] * 2,
            collide=[1] * n_agents + [0, 0],
            type=[AGENT] * n_agents + [MARKER, MARKER],
        )
        kw.update(overrides)
        return WorldConfig(**kw)
