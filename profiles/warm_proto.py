"""Prototype of the warm-started assignment (env_team_kernel, fp32 production mode):
   state kept across steps: column duals v, matching pi (c4r / r4c), valid flag.
   step: (0) invalid -> cold scipy-order solve
         (1) R rounds of Bellman-Ford re-centring of v under the OLD matching
         (2) u = row minima of C - v; a row keeps its match iff its matched edge attains the minimum
         (3) Dijkstra augmentation of the free rows (any order), dual updates as in the cold solver
         (4) certificate: the graph of small-slack (<= tol) non-matching edges, as a row graph, must be acyclic
             -> the optimum is unique by a margin -> equal to scipy's; else cold solve.
"""
import numpy as np, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from scipy.optimize import linear_sum_assignment

def costs(cfg, o, name, N):
    """SPEC §5 cost matrices [B, agent, slot] of an oracle env batch."""
    pos = o.agent_state[..., :2].astype(np.float64)
    lm = o.landmark_pos.astype(np.float64)
    st = np.asarray(cfg.slot_table, np.float64).reshape(N, 2)
    if name == "polygon":
        slots = lm[:, 0:1, :] + cfg.polygon_radius * st[None]
    else:
        slots = lm[:, 0:1, :] + st[None, :, 0:1] * (lm[:, 1:2, :] - lm[:, 0:1, :])
    return np.sqrt(((slots[:, None, :, :] - pos[:, :, None, :]) ** 2).sum(-1))


def cold(C):
    r, c = linear_sum_assignment(C)
    return c

def dijkstra_augment(C, u, v, c4r, r4c, cur):
    N = C.shape[0]
    spc = np.full(N, np.inf, C.dtype); path = -np.ones(N, int)
    SR = np.zeros(N, bool); SC = np.zeros(N, bool)
    i = cur; minval = C.dtype.type(0); sink = -1
    while sink < 0:
        SR[i] = True
        r = (minval + C[i] - u[i] - v).astype(C.dtype)
        upd = (~SC) & (r < spc)
        spc[upd] = r[upd]; path[upd] = i
        cand = np.where(~SC, spc, np.inf)
        j = int(np.argmin(cand)); minval = cand[j]
        if not np.isfinite(minval): return False
        SC[j] = True
        if r4c[j] < 0: sink = j
        else: i = r4c[j]
    u[cur] += minval
    for r_ in range(N):
        if SR[r_] and r_ != cur: u[r_] += minval - spc[c4r[r_]]
    v[SC] -= (minval - spc[SC])
    j = sink
    while True:
        i = path[j]; r4c[j] = i
        j, c4r[i] = c4r[i], j
        if i == cur: break
    return True

class Warm:
    def __init__(self, N, dtype, tol, R=2):
        self.N, self.dt, self.tol, self.R = N, dtype, dtype(tol), R
        self.valid = False
        self.stats = dict(cold=0, warm=0, cert_fail=0, free=[])
    def solve(self, C):
        N = self.N; C = C.astype(self.dt)
        if not self.valid:
            return self._cold(C)
        v, c4r, r4c = self.v.copy(), self.c4r.copy(), self.r4c.copy()
        # (1) BF re-centring: v_k <- min(v_k, min_r v_pi(r) + C[r][k] - C[r][pi(r)])
        for _ in range(self.R):
            w = v[c4r] - C[np.arange(N), c4r]              # per row r
            v = np.minimum(v, (w[:, None] + C).min(0)).astype(self.dt)
        # (2) row minima, kept matches
        red = (C - v[None, :]).astype(self.dt)
        u = red.min(1)
        free = []
        for r in range(N):
            k = c4r[r]
            if not (red[r, k] == u[r]):
                free.append(r); r4c[k] = -1; c4r[r] = -1
        self.stats['free'].append(len(free))
        # (3) augment
        for cur in free:
            if not dijkstra_augment(C, u, v, c4r, r4c, cur):
                return self._cold(C)
        # (4) certificate
        slack = (C - u[:, None] - v[None, :]).astype(self.dt)
        small = slack <= self.tol
        M = np.zeros((N, N), bool)                         # row graph: r -> r4c[k] for small non-matching (r, k)
        for r in range(N):
            for k in range(N):
                if small[r, k] and c4r[r] != k: M[r, r4c[k]] = True
        alive = np.ones(N, bool)
        changed = True
        while changed:
            changed = False
            for r in range(N):
                if alive[r] and not (M[r] & alive).any():
                    alive[r] = False; changed = True
        if alive.any():
            self.stats['cert_fail'] += 1
            return self._cold(C)
        self.stats['warm'] += 1
        self.v, self.c4r, self.r4c = v, c4r, r4c
        return c4r.copy()
    def _cold(self, C):
        self.stats['cold'] += 1
        N = self.N
        # scipy-order cold solve with duals (same routine, rows in order, from zero duals)
        u = np.zeros(N, self.dt); v = np.zeros(N, self.dt)
        c4r = -np.ones(N, int); r4c = -np.ones(N, int)
        for cur in range(N):
            dijkstra_augment(C, u, v, c4r, r4c, cur)
        ref = cold(C)
        self.v, self.c4r, self.r4c, self.valid = v, ref.copy(), np.argsort(ref), True
        # NOTE: the kernel's cold path IS the scipy-order solver; here scipy gives the permutation and the
        # prototype's own Dijkstra the duals (they belong to an optimal matching; if it is not scipy's, the
        # next warm step re-derives feasibility anyway)
        if not (c4r == ref).all():
            self.v = v  # duals stay valid for any optimal matching
        return ref

if __name__ == "__main__" and len(sys.argv) == 1:
    from tests._util import make_cfg, random_actions
    from oracle import gsm_oracle as O
    for dtype, tol in ((np.float32, 1e-4), (np.float64, 1e-9)):
        for name, N in (("polygon", 12), ("line", 12), ("polygon", 6), ("line", 5), ("polygon", 3)):
            B = 48
            cfg = make_cfg(name, N, "f64")
            o = O.OracleEnv(cfg, B); o.reset(7)
            # tie-heavy envs: agents exactly on the slots in reversed order, zero velocity, first action 0
            rng = np.random.default_rng(3)
            ws = [Warm(N, dtype, tol) for _ in range(B)]
            bad = 0; tot = 0
            for t in range(40):
                a = random_actions(cfg, rng, (B,))
                if t < 2: a[:] = 0
                o.step(a)
                C = costs(cfg, o, name, N)
                for b in range(B):
                    Cb = C[b].astype(dtype)
                    if b < 8: Cb = np.round(Cb * 4) / 4          # quantised costs: many exact ties
                    got = ws[b].solve(Cb)
                    ref = cold(Cb)
                    tot += 1; bad += int(not (got == ref).all())
            st = {k: sum(w.stats[k] for w in ws) for k in ("cold", "warm", "cert_fail")}
            fr = np.concatenate([w.stats['free'] for w in ws])
            print(dtype.__name__, name, N, "mismatch", bad, "/", tot, st, "mean free", round(fr.mean(), 2))

def sweep_tol():
    from tests._util import make_cfg, random_actions
    from oracle import gsm_oracle as O
    for tol in (1e-4, 3e-5, 1e-5, 3e-6, 1e-6):
        name, N, B = "polygon", 12, 64
        cfg = make_cfg(name, N, "f64")
        o = O.OracleEnv(cfg, B); o.reset(11)
        rng = np.random.default_rng(5)
        ws = [Warm(N, np.float32, tol) for _ in range(B)]
        bad = tot = 0
        for t in range(60):
            o.step(random_actions(cfg, rng, (B,)))
            C = costs(cfg, o, name, N)
            for b in range(B):
                Cb = C[b].astype(np.float32)
                got = ws[b].solve(Cb); ref = cold(Cb.astype(np.float64))
                tot += 1; bad += int(not (got == ref).all())
        st = {k: sum(w.stats[k] for w in ws) for k in ("cold", "warm", "cert_fail")}
        print("tol", tol, "mismatch", bad, "/", tot, st)
if len(sys.argv) > 1 and sys.argv[1] == "tol":
    sweep_tol()
