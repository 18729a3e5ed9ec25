"""Writes a SYNTHETIC `gsmarl/` source tree shaped like GSMARL.egg-info/SOURCES.txt:6-35, for one
purpose only: proving tools/unblock.py end to end while the real sources are withheld
(reference readme.md:1).

    python tools/fake_gsmarl.py /tmp/fake_ref [--damping 0.5]

The tree is NOT GS-MARL and nothing in the product or the oracle reads it.  It is a stand-in
written in the MPE / InforMARL lineage's *style* (per-world Python objects, `gym.spaces`,
scenario callbacks, an argparse `get_config()`), whose arithmetic is SPEC.md — so that running the
kit on it must end in an all-PASS diff table, and running it on a tree with one constant or one
formula changed (`--damping`, `--contact-margin`, `--reward-bug`) must flag exactly the SPEC
sections that changed (tests/test_unblock_kit.py).
"""
from __future__ import annotations

import argparse
import os
import textwrap

FILES = {}

FILES["setup.py"] = '''
from setuptools import setup, find_packages
setup(name="GSMARL", version="0.0.0-synthetic", packages=find_packages())
'''

FILES["GSMARL.egg-info/SOURCES.txt"] = "\n".join("""setup.py
GSMARL.egg-info/SOURCES.txt
gsmarl/__init__.py
gsmarl/config.py
gsmarl/envs/__init__.py
gsmarl/envs/mpe_env/__init__.py
gsmarl/envs/mpe_env/env_wrappers.py
gsmarl/envs/mpe_env/make_env.py
gsmarl/envs/mpe_env/multiagent/__init__.py
gsmarl/envs/mpe_env/multiagent/core.py
gsmarl/envs/mpe_env/multiagent/environment.py
gsmarl/envs/mpe_env/multiagent/scenario.py
gsmarl/envs/mpe_env/multiagent/scenarios/__init__.py
gsmarl/envs/mpe_env/multiagent/scenarios/exp1.py
gsmarl/envs/mpe_env/multiagent/scenarios/simple_formation.py
gsmarl/envs/mpe_env/multiagent/scenarios/simple_line.py""".split("\n")) + "\n"

FILES["gsmarl/__init__.py"] = '__version__ = "0.0.0-synthetic"\n'
FILES["gsmarl/envs/__init__.py"] = ""
FILES["gsmarl/envs/mpe_env/__init__.py"] = ""
FILES["gsmarl/envs/mpe_env/multiagent/__init__.py"] = ""

FILES["gsmarl/config.py"] = '''
import argparse


def get_config():
    parser = argparse.ArgumentParser(description="synthetic gsmarl config")
    parser.add_argument("--scenario_name", type=str, default="exp1")
    parser.add_argument("--num_agents", type=int, default=3)
    parser.add_argument("--num_obstacles", type=int, default=None)
    parser.add_argument("--episode_length", type=int, default=None)
    parser.add_argument("--n_rollout_threads", type=int, default=4)
    parser.add_argument("--max_edge_dist", type=float, default=1.0)
    parser.add_argument("--max_nbrs", type=int, default=None)
    parser.add_argument("--world_size", type=float, default=None)
    parser.add_argument("--seed", type=int, default=1)
    return parser
'''

FILES["gsmarl/envs/mpe_env/multiagent/core.py"] = '''
import numpy as np


class EntityState(object):
    def __init__(self):
        self.p_pos = None
        self.p_vel = None


class Action(object):
    def __init__(self):
        self.u = None


class Entity(object):
    def __init__(self):
        self.name = ""
        self.size = 0.050
        self.movable = False
        self.collide = True
        self.max_speed = None
        self.accel = None
        self.state = EntityState()
        self.initial_mass = 1.0

    @property
    def mass(self):
        return self.initial_mass


class Landmark(Entity):
    def __init__(self):
        super(Landmark, self).__init__()
        self.kind = "landmark"      # goal | obstacle | marker


class Agent(Entity):
    def __init__(self):
        super(Agent, self).__init__()
        self.movable = True
        self.action = Action()
        self.goal = None


class World(object):
    def __init__(self):
        self.agents = []
        self.landmarks = []
        self.dim_p = 2
        self.dt = 0.1
        self.damping = @DAMPING@
        self.contact_force = 1e+2
        self.contact_margin = @CONTACT_MARGIN@
        self.max_edge_dist = 1.0
        self.world_size = 1.0
        self.current_time_step = 0

    @property
    def entities(self):
        return self.agents + self.landmarks

    def step(self):
        p_force = [None] * len(self.agents)
        for i, agent in enumerate(self.agents):
            p_force[i] = np.array(agent.action.u, dtype=np.float64)
        # environment force: per agent, other entities in ascending index
        for i, a in enumerate(self.agents):
            if not a.collide:
                continue
            for b in self.entities:
                if b is a or not b.collide:
                    continue
                delta = a.state.p_pos - b.state.p_pos
                dist = np.sqrt(delta[0] * delta[0] + delta[1] * delta[1])
                dist_min = a.size + b.size
                k = self.contact_margin
                penetration = np.logaddexp(0, -(dist - dist_min) / k) * k
                p_force[i] = p_force[i] + self.contact_force * delta / dist * penetration
        for i, a in enumerate(self.agents):
            v_old = a.state.p_vel
            a.state.p_vel = a.state.p_vel * (1 - self.damping)
            a.state.p_vel = a.state.p_vel + (p_force[i] / a.mass) * self.dt
            if a.max_speed is not None:
                speed = np.sqrt(np.square(a.state.p_vel[0]) + np.square(a.state.p_vel[1]))
                if speed > a.max_speed:
                    a.state.p_vel = a.state.p_vel / speed * a.max_speed
            a.state.p_pos = a.state.p_pos + @POS_VEL@ * self.dt
        self.current_time_step += 1
'''

FILES["gsmarl/envs/mpe_env/multiagent/scenario.py"] = '''
class BaseScenario(object):
    def make_world(self, args):
        raise NotImplementedError()

    def reset_world(self, world):
        raise NotImplementedError()
'''

FILES["gsmarl/envs/mpe_env/multiagent/scenarios/__init__.py"] = '''
import importlib.util
import os.path as osp


def load(name):
    pathname = osp.join(osp.dirname(__file__), name)
    spec = importlib.util.spec_from_file_location(name[:-3], pathname)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
'''

_SCENARIO_COMMON = '''
import numpy as np
from scipy.optimize import linear_sum_assignment
from gsmarl.envs.mpe_env.multiagent.core import World, Agent, Landmark
from gsmarl.envs.mpe_env.multiagent.scenario import BaseScenario

ENTITY_TYPE = {"agent": 0, "goal": 1, "obstacle": 2, "marker": 3}


def _new_agent(i):
    a = Agent()
    a.name = "agent %d" % i
    a.size = 0.10
    a.accel = 5.0
    a.max_speed = 1.3
    a.initial_mass = 1.0
    a.collide = True
    return a


class GraphMixin(object):
    """observation / graph / reward / cost callbacks shared by the synthetic scenarios."""

    def targets(self, world):
        raise NotImplementedError()

    def observation(self, agent, world):
        i = world.agents.index(agent)
        tgt = self.targets(world)[1][i]
        return np.concatenate([agent.state.p_vel, agent.state.p_pos, tgt - agent.state.p_pos])

    def graph_observation(self, agent, world):
        """node_obs [K][6] (zero padded), adj row over all entities (0/1), neighbour ids [K] (-1 padded)."""
        i = world.agents.index(agent)
        K = world.max_nbrs
        ents = world.entities
        node = np.zeros((K, 6))
        ids = -np.ones(K, dtype=np.int64)
        adj = np.zeros(len(ents), dtype=np.int64)
        cnt = 0
        for e, b in enumerate(ents):
            if b is agent:
                continue
            d = b.state.p_pos - agent.state.p_pos
            dist = np.sqrt(d[0] * d[0] + d[1] * d[1])
            nb = dist < world.max_edge_dist or (self.own_goal_always and e == len(world.agents) + i)
            if nb:
                adj[e] = 1
                if cnt < K:
                    vel = b.state.p_vel if b.state.p_vel is not None else np.zeros(2)
                    node[cnt] = [d[0], d[1], vel[0] - agent.state.p_vel[0], vel[1] - agent.state.p_vel[1],
                                 dist, float(ENTITY_TYPE[getattr(b, "kind", "agent")])]
                    ids[cnt] = e
                    cnt += 1
        return node, adj, ids

    def reward(self, agent, world):
        i = world.agents.index(agent)
        tgt = self.targets(world)[1][i]
        g = tgt - agent.state.p_pos
        d = np.sqrt(g[0] * g[0] + g[1] * g[1])
        return @REWARD_EXPR@

    def cost(self, agent, world):
        n = 0
        for b in world.entities:
            if b is agent:
                continue
            is_agent = b in world.agents
            if not (is_agent or (self.cost_obstacles and getattr(b, "kind", "") == "obstacle")):
                continue
            d = b.state.p_pos - agent.state.p_pos
            if np.sqrt(d[0] * d[0] + d[1] * d[1]) < agent.size + b.size:
                n += 1
        return float(n)

    def info(self, agent, world):
        i = world.agents.index(agent)
        return {"assign": int(self.targets(world)[0][i])}
'''

FILES["gsmarl/envs/mpe_env/multiagent/scenarios/exp1.py"] = _SCENARIO_COMMON + '''

class Scenario(GraphMixin, BaseScenario):
    own_goal_always = True
    cost_obstacles = True

    def make_world(self, args):
        world = World()
        n = args.num_agents
        n_obs = n if getattr(args, "num_obstacles", None) is None else args.num_obstacles
        world.world_size = np.sqrt(max(n, 3) / 3.0) if getattr(args, "world_size", None) is None else args.world_size
        world.max_edge_dist = args.max_edge_dist
        E = 2 * n + n_obs
        world.max_nbrs = args.max_nbrs if getattr(args, "max_nbrs", None) else (E - 1 if E - 1 <= 8 else min(32, (E - 1) // 4 * 4))
        world.episode_length = args.episode_length or 25
        world.agents = [_new_agent(i) for i in range(n)]
        world.landmarks = []
        for i in range(n):
            g = Landmark(); g.kind = "goal"; g.name = "goal %d" % i; g.size = 0.05; g.collide = False
            world.landmarks.append(g)
        for i in range(n_obs):
            o = Landmark(); o.kind = "obstacle"; o.name = "obstacle %d" % i; o.size = 0.16; o.collide = True
            world.landmarks.append(o)
        self.reset_world(world)
        return world

    def reset_world(self, world):
        s = world.world_size
        for e in world.entities:
            e.state.p_pos = np.random.uniform(-s, +s, world.dim_p)
            e.state.p_vel = np.zeros(world.dim_p)
        world.current_time_step = 0

    def targets(self, world):
        n = len(world.agents)
        return list(range(n)), [world.landmarks[i].state.p_pos for i in range(n)]
'''

FILES["gsmarl/envs/mpe_env/multiagent/scenarios/simple_formation.py"] = _SCENARIO_COMMON + '''

class Scenario(GraphMixin, BaseScenario):
    own_goal_always = False
    cost_obstacles = False
    target_radius = 0.5

    def make_world(self, args):
        world = World()
        n = args.num_agents
        world.world_size = np.sqrt(max(n, 3) / 3.0)
        world.max_edge_dist = args.max_edge_dist
        world.max_nbrs = args.max_nbrs if getattr(args, "max_nbrs", None) else (n if n <= 8 else min(32, n // 4 * 4))
        world.episode_length = args.episode_length or 100      # reference readme.md:101
        world.agents = [_new_agent(i) for i in range(n)]
        m = Landmark(); m.kind = "marker"; m.name = "landmark 0"; m.size = 0.16; m.collide = False
        world.landmarks = [m]
        self.reset_world(world)
        return world

    def reset_world(self, world):
        s = world.world_size
        for a in world.agents:
            a.state.p_pos = np.random.uniform(-s, +s, world.dim_p)
            a.state.p_vel = np.zeros(world.dim_p)
        for l in world.landmarks:
            l.state.p_pos = np.random.uniform(-0.5 * s, +0.5 * s, world.dim_p)
            l.state.p_vel = np.zeros(world.dim_p)
        world.current_time_step = 0

    def slots(self, world):
        n = len(world.agents)
        c = world.landmarks[0].state.p_pos
        return [np.array([c[0] + self.target_radius * np.cos(2.0 * np.pi * k / n),
                          c[1] + self.target_radius * np.sin(2.0 * np.pi * k / n)]) for k in range(n)]

    def targets(self, world):
        slots = self.slots(world)
        n = len(world.agents)
        cost = np.zeros((n, n))
        for i, a in enumerate(world.agents):
            for k in range(n):
                d = slots[k] - a.state.p_pos
                cost[i, k] = np.sqrt(d[0] * d[0] + d[1] * d[1])
        _, cols = linear_sum_assignment(cost)
        return [int(c) for c in cols], [slots[int(cols[i])] for i in range(n)]
'''

FILES["gsmarl/envs/mpe_env/multiagent/scenarios/simple_line.py"] = _SCENARIO_COMMON + '''

class Scenario(GraphMixin, BaseScenario):
    own_goal_always = False
    cost_obstacles = False

    def make_world(self, args):
        world = World()
        n = args.num_agents
        world.world_size = np.sqrt(max(n, 3) / 3.0)
        world.max_edge_dist = args.max_edge_dist
        E = n + 2
        world.max_nbrs = args.max_nbrs if getattr(args, "max_nbrs", None) else (E - 1 if E - 1 <= 8 else min(32, (E - 1) // 4 * 4))
        world.episode_length = args.episode_length or 100      # reference readme.md:101
        world.agents = [_new_agent(i) for i in range(n)]
        world.landmarks = []
        for k in range(2):
            m = Landmark(); m.kind = "marker"; m.name = "landmark %d" % k; m.size = 0.16; m.collide = False
            world.landmarks.append(m)
        self.reset_world(world)
        return world

    def reset_world(self, world):
        s = world.world_size
        for e in world.entities:
            e.state.p_pos = np.random.uniform(-s, +s, world.dim_p)
            e.state.p_vel = np.zeros(world.dim_p)
        world.current_time_step = 0

    def targets(self, world):
        n = len(world.agents)
        a_, b_ = world.landmarks[0].state.p_pos, world.landmarks[1].state.p_pos
        slots = []
        for k in range(n):
            f = (k + 1.0) / (n + 1.0)
            slots.append(np.array([a_[0] + f * (b_[0] - a_[0]), a_[1] + f * (b_[1] - a_[1])]))
        cost = np.zeros((n, n))
        for i, a in enumerate(world.agents):
            for k in range(n):
                d = slots[k] - a.state.p_pos
                cost[i, k] = np.sqrt(d[0] * d[0] + d[1] * d[1])
        _, cols = linear_sum_assignment(cost)
        return [int(c) for c in cols], [slots[int(cols[i])] for i in range(n)]
'''

FILES["gsmarl/envs/mpe_env/multiagent/environment.py"] = '''
import gym
from gym import spaces
import numpy as np


class MultiAgentEnv(gym.Env):
    """Fixed-size observations, no cost."""
    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, world, reset_callback=None, reward_callback=None, observation_callback=None,
                 info_callback=None, done_callback=None, discrete_action=True):
        self.world = world
        self.agents = self.world.agents
        self.n = len(world.agents)
        self.reset_callback = reset_callback
        self.reward_callback = reward_callback
        self.observation_callback = observation_callback
        self.info_callback = info_callback
        self.done_callback = done_callback
        self.discrete_action_space = discrete_action
        self.action_space = []
        self.observation_space = []
        self.share_observation_space = []
        for agent in self.agents:
            self.action_space.append(spaces.Discrete(world.dim_p * 2 + 1))
            obs_dim = len(observation_callback(agent, self.world))
            self.observation_space.append(spaces.Box(low=-np.inf, high=+np.inf, shape=(obs_dim,), dtype=np.float32))
        share = sum(s.shape[0] for s in self.observation_space)
        self.share_observation_space = [spaces.Box(low=-np.inf, high=+np.inf, shape=(share,), dtype=np.float32)
                                        for _ in range(self.n)]

    def seed(self, seed=None):
        np.random.seed(1 if seed is None else seed)

    def _set_action(self, action, agent):
        agent.action.u = np.zeros(self.world.dim_p)
        action = np.asarray(action)
        if action.ndim == 0 or action.size == 1:          # index
            onehot = np.zeros(self.world.dim_p * 2 + 1)
            onehot[int(action)] = 1.0
            action = onehot
        agent.action.u[0] += action[1] - action[2]
        agent.action.u[1] += action[3] - action[4]
        sensitivity = 5.0 if agent.accel is None else agent.accel
        agent.action.u *= sensitivity

    def _get_obs(self, agent):
        return self.observation_callback(agent, self.world)

    def _get_reward(self, agent):
        return self.reward_callback(agent, self.world)

    def _get_done(self, agent):
        return self.world.current_time_step >= self.world.episode_length

    def _get_info(self, agent):
        return {} if self.info_callback is None else self.info_callback(agent, self.world)

    def step(self, action_n):
        for i, agent in enumerate(self.agents):
            self._set_action(action_n[i], agent)
        self.world.step()
        obs_n = [self._get_obs(a) for a in self.agents]
        reward_n = [[self._get_reward(a)] for a in self.agents]
        done_n = [self._get_done(a) for a in self.agents]
        info_n = [self._get_info(a) for a in self.agents]
        return obs_n, reward_n, done_n, info_n

    def reset(self):
        self.reset_callback(self.world)
        return [self._get_obs(a) for a in self.agents]


class MultiAgentConstrainEnv(MultiAgentEnv):
    """Fixed-size observations + cost."""

    def __init__(self, world, reset_callback=None, reward_callback=None, observation_callback=None,
                 info_callback=None, done_callback=None, cost_callback=None, discrete_action=True):
        super(MultiAgentConstrainEnv, self).__init__(world, reset_callback, reward_callback, observation_callback,
                                                     info_callback, done_callback, discrete_action)
        self.cost_callback = cost_callback

    def _get_cost(self, agent):
        return self.cost_callback(agent, self.world)

    def step(self, action_n):
        obs_n, reward_n, done_n, info_n = super(MultiAgentConstrainEnv, self).step(action_n)
        cost_n = [[self._get_cost(a)] for a in self.agents]
        return obs_n, reward_n, cost_n, done_n, info_n


class MultiAgentGraphConstrainEnv(MultiAgentConstrainEnv):
    """Variable-size graph observations + cost."""

    def __init__(self, world, reset_callback=None, reward_callback=None, observation_callback=None,
                 graph_observation_callback=None, info_callback=None, done_callback=None, cost_callback=None,
                 discrete_action=True):
        super(MultiAgentGraphConstrainEnv, self).__init__(world, reset_callback, reward_callback,
                                                          observation_callback, info_callback, done_callback,
                                                          cost_callback, discrete_action)
        self.graph_observation_callback = graph_observation_callback
        node, adj, _ = graph_observation_callback(self.agents[0], world)
        self.node_observation_space = [spaces.Box(low=-np.inf, high=+np.inf, shape=node.shape, dtype=np.float32)
                                       for _ in range(self.n)]
        self.adj_observation_space = [spaces.Box(low=0, high=1, shape=adj.shape, dtype=np.float32)
                                      for _ in range(self.n)]

    def _graph(self):
        g = [self.graph_observation_callback(a, self.world) for a in self.agents]
        return [x[0] for x in g], [x[1] for x in g], [x[2] for x in g]

    def step(self, action_n):
        obs_n, reward_n, cost_n, done_n, info_n = super(MultiAgentGraphConstrainEnv, self).step(action_n)
        node_n, adj_n, ids_n = self._graph()
        for info, ids in zip(info_n, ids_n):
            info["nbr_idx"] = ids
        return obs_n, (node_n, adj_n), reward_n, cost_n, done_n, info_n

    def reset(self):
        obs_n = super(MultiAgentGraphConstrainEnv, self).reset()
        node_n, adj_n, _ = self._graph()
        return obs_n, (node_n, adj_n)
'''

FILES["gsmarl/envs/mpe_env/make_env.py"] = '''
from gsmarl.envs.mpe_env.multiagent.environment import MultiAgentGraphConstrainEnv
import gsmarl.envs.mpe_env.multiagent.scenarios as scenarios


def MPEEnv(args):
    scenario = scenarios.load(args.scenario_name + ".py").Scenario()
    world = scenario.make_world(args)
    return MultiAgentGraphConstrainEnv(world, reset_callback=scenario.reset_world, reward_callback=scenario.reward,
                                       observation_callback=scenario.observation,
                                       graph_observation_callback=scenario.graph_observation,
                                       info_callback=scenario.info, cost_callback=scenario.cost)
'''

FILES["gsmarl/envs/mpe_env/env_wrappers.py"] = '''
import numpy as np


class DummyVecEnv(object):
    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)

    def reset(self):
        return [e.reset() for e in self.envs]

    def step(self, actions):
        return [e.step(a) for e, a in zip(self.envs, actions)]
'''


def write_tree(root: str, damping: float = 0.25, contact_margin: float = 1e-3, reward_bug: bool = False,
               physics_bug: bool = False) -> str:
    """Write the synthetic tree under `root` (created); returns root."""
    reward_expr = "(0.0 - 1.0 * d) + (1.0 if d < 0.1 else 0.0)"
    if reward_bug:
        reward_expr = "(0.0 - 1.0 * d * d) + (1.0 if d < 0.1 else 0.0)"     # squared distance: not SPEC §7
    for rel, body in FILES.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        body = textwrap.dedent(body).lstrip("\n")
        body = (body.replace("@DAMPING@", repr(float(damping))).replace("@CONTACT_MARGIN@", repr(float(contact_margin)))
                .replace("@REWARD_EXPR@", reward_expr)
                .replace("@POS_VEL@", "v_old" if physics_bug else "a.state.p_vel"))     # explicit Euler: not SPEC §4
        with open(path, "w") as f:
            f.write(body)
    with open(os.path.join(root, "readme.md"), "w") as f:
        f.write("# SYNTHETIC stand-in tree written by tools/fake_gsmarl.py - NOT GS-MARL\n")
    return root


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("root")
    ap.add_argument("--damping", type=float, default=0.25)
    ap.add_argument("--contact-margin", type=float, default=1e-3)
    ap.add_argument("--reward-bug", action="store_true")
    ap.add_argument("--physics-bug", action="store_true")
    a = ap.parse_args()
    print(write_tree(a.root, a.damping, a.contact_margin, a.reward_bug, a.physics_bug))
