"""Second, independent restatement of SPEC.md in the reference's *style*: one Python
object per world, per-agent Python loops, numpy scalar arithmetic, `np.logaddexp` and
`scipy.optimize.linear_sum_assignment` called directly — i.e. how a numpy MPE-family env
(environment.py / core.py / scenarios/*.py, GSMARL.egg-info/SOURCES.txt:14,15,21-25) runs
on the CPU.  Used (a) to cross-check the C oracle, (b) as the "reference-style" CPU
baseline leg of bench.py (one env per object, `n_rollout_threads` subprocess workers).

TEST INFRASTRUCTURE ONLY — PARITY UNPINNED.  It is written from SPEC.md, not from the
withheld GS-MARL sources and not from upstream MPE code.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment

from . import worlds


class Entity:
    def __init__(self, idx, etype, size, collide):
        self.idx, self.type, self.size, self.collide = idx, etype, size, collide
        self.p_pos = np.zeros(2)
        self.p_vel = np.zeros(2)


class World:
    """SPEC §2-4."""

    def __init__(self, cfg):
        self.cfg = cfg
        R = cfg.np_real
        self.R = R
        N = cfg.n_agents
        self.entities = [Entity(e, cfg.type[e], R(cfg.size[e]), bool(cfg.collide[e]))
                         for e in range(cfg.n_entities)]
        self.agents, self.landmarks = self.entities[:N], self.entities[N:]
        for e in self.entities:
            e.p_pos = np.zeros(2, R)
            e.p_vel = np.zeros(2, R)
        self.t = 0

    def step(self, controls):
        c, R = self.cfg, self.R
        cf, km, dt, damp = R(c.contact_force), R(c.contact_margin), R(c.dt), R(c.damping)
        new = []
        for i, a in enumerate(self.agents):
            f = R(c.accel[i]) * controls[i]
            fx, fy = f[0], f[1]
            if a.collide:
                for b in self.entities:
                    if b is a or not b.collide:
                        continue
                    dx, dy = a.p_pos[0] - b.p_pos[0], a.p_pos[1] - b.p_pos[1]
                    dist = np.sqrt(dx * dx + dy * dy)
                    x = -(dist - (a.size + b.size)) / km
                    pen = np.logaddexp(R(0), x) * km
                    fx = fx + cf * dx / dist * pen
                    fy = fy + cf * dy / dist * pen
            vx, vy = a.p_vel[0] * (R(1) - damp), a.p_vel[1] * (R(1) - damp)
            m = R(c.mass[i])
            vx, vy = vx + (fx / m) * dt, vy + (fy / m) * dt
            ms = R(c.max_speed[i])
            if ms > 0:
                sp = np.sqrt(vx * vx + vy * vy)
                if sp > ms:
                    vx, vy = vx / sp * ms, vy / sp * ms
            new.append((a.p_pos[0] + vx * dt, a.p_pos[1] + vy * dt, vx, vy))
        for a, (px, py, vx, vy) in zip(self.agents, new):
            a.p_pos = np.array([px, py], R)
            a.p_vel = np.array([vx, vy], R)
        self.t += 1


class PyEnv:
    """One env; step(actions) -> dict of per-agent arrays (same names as gsm_step_io)."""

    def __init__(self, cfg):
        cfg = worlds.as_world(cfg)        # shapes / slot tables re-derived from SPEC.md, not taken from the caller
        self.cfg, self.world = cfg, World(cfg)

    def set_state(self, agent_state, landmark_pos, t=0):
        for a, s in zip(self.world.agents, agent_state):
            a.p_pos = np.array(s[:2], self.world.R)
            a.p_vel = np.array(s[2:], self.world.R)
        for l, p in zip(self.world.landmarks, landmark_pos):
            l.p_pos = np.array(p, self.world.R)
        self.world.t = int(t)

    def get_state(self):
        return np.array([[*a.p_pos, *a.p_vel] for a in self.world.agents], self.world.R)

    # -- scenario callbacks (SPEC §5-7) ----------------------------------------------
    def _targets(self):
        c, w, R = self.cfg, self.world, self.world.R
        N = c.n_agents
        if c.scenario == "navigation":
            return list(range(N)), [w.landmarks[i].p_pos for i in range(N)]
        lm = w.landmarks
        slots = []
        for k in range(N):
            if c.scenario == "polygon":
                rad = R(c.polygon_radius)
                slots.append(np.array([lm[0].p_pos[0] + rad * R(c.slot_table[k][0]),
                                       lm[0].p_pos[1] + rad * R(c.slot_table[k][1])], R))
            else:
                f = R(c.slot_table[k][0])
                slots.append(np.array([lm[0].p_pos[0] + f * (lm[1].p_pos[0] - lm[0].p_pos[0]),
                                       lm[0].p_pos[1] + f * (lm[1].p_pos[1] - lm[0].p_pos[1])], R))
        cost = np.zeros((N, N), R)
        for i, a in enumerate(w.agents):
            for k in range(N):
                dx, dy = slots[k][0] - a.p_pos[0], slots[k][1] - a.p_pos[1]
                cost[i, k] = np.sqrt(dx * dx + dy * dy)
        rows, cols = linear_sum_assignment(cost)
        assign = [int(cols[i]) for i in range(N)]
        return assign, [slots[assign[i]] for i in range(N)]

    def observe(self, with_rcd=True):
        c, w, R = self.cfg, self.world, self.world.R
        N, E, K = c.n_agents, c.n_entities, c.max_nbrs
        out = {k: np.zeros(s[1:], d) for k, (d, s) in c.io_shapes(1).items() if k != "actions"}
        out["nbr_idx"][...] = -1
        assign, targets = self._targets()
        rew = []
        for i, a in enumerate(w.agents):
            tg = targets[i]
            out["obs"][i] = [a.p_vel[0], a.p_vel[1], a.p_pos[0], a.p_pos[1],
                             tg[0] - a.p_pos[0], tg[1] - a.p_pos[1]]
            out["assign"][i] = assign[i]
            cnt = ncol = 0
            for b in w.entities:
                if b is a:
                    continue
                dx, dy = b.p_pos[0] - a.p_pos[0], b.p_pos[1] - a.p_pos[1]
                dist = np.sqrt(dx * dx + dy * dy)
                nb = bool(dist < R(c.sensing_radius))
                if c.own_goal_always and c.scenario == "navigation" and b.idx == N + i:
                    nb = True
                if nb:
                    out["adj"][i, b.idx >> 5] |= np.uint32(1 << (b.idx & 31))
                    if cnt < K:
                        out["nbr_idx"][i, cnt] = b.idx
                        out["nbr_feat"][i, cnt] = [dx, dy, b.p_vel[0] - a.p_vel[0],
                                                   b.p_vel[1] - a.p_vel[1], dist, R(b.type)]
                        cnt += 1
                if dist < a.size + b.size:
                    if b.idx < N or (c.cost_obstacles and b.type == worlds.ENT_OBSTACLE):
                        ncol += 1
            out["nbr_cnt"][i] = cnt
            gx, gy = tg[0] - a.p_pos[0], tg[1] - a.p_pos[1]
            d = np.sqrt(gx * gx + gy * gy)
            rew.append((R(0) - R(c.w_dist) * d) + (R(c.w_goal) if d < R(c.goal_tol) else R(0)))
            out["cost"][i] = ncol
            out["done"][i] = w.t >= c.episode_length
        if c.share_reward:
            s = rew[0]
            for r in rew[1:]:
                s = s + r
            rew = [s / R(N)] * N
        out["reward"][:] = rew
        if not with_rcd:
            for k in ("reward", "cost", "done"):
                out[k][...] = 0
        return out

    def step(self, actions):
        c, R = self.cfg, self.world.R
        ctrl = []
        for i in range(c.n_agents):
            if c.action_mode == "discrete":
                a = int(actions[i])
                u = c.discrete_u[a] if 0 <= a < len(c.discrete_u) else (0.0, 0.0)
                ctrl.append(np.array(u, R))
            else:
                ctrl.append(np.array(actions[i], R))
        self.world.step(ctrl)
        return self.observe()
