"""Generates the committed fixtures under tests/golden/.  Run from the repo root:
    python tests/golden/make_golden.py

lsa_scipy.npz      cost matrices (random, tie-heavy integer, constant, degenerate
                   geometry) with the assignment returned by THIS image's
                   scipy.optimize.linear_sum_assignment (version stored in the file).
                   This is the one external pin of the oracle: the reference pins
                   scipy==1.7.3 (requirements.txt:101), same rectangular_lsap algorithm.
traj_*.npz         25-step trajectories of the declared model (SPEC.md) produced by the
                   pure-Python restatement oracle/py_env.py (np.logaddexp + scipy LSA).
                   The reference ships no golden vectors and its env is withheld, so
                   these pin the C oracle and the CUDA kernels to SPEC.md, NOT to GS-MARL.
"""
import os
import zlib
import sys

import numpy as np
import scipy
from scipy.optimize import linear_sum_assignment

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gsm_oracle as O, py_env, worlds  # noqa: E402   (oracle/ only: no product import)

HERE = os.path.dirname(os.path.abspath(__file__))


def lsa_cases():
    rng = np.random.default_rng(20261018)
    mats = []
    for n in range(1, 17):
        for _ in range(12):
            mats.append(rng.random((n, n)))
        for hi in (2, 3, 5):
            for _ in range(8):
                mats.append(rng.integers(0, hi, (n, n)).astype(np.float64))
        mats.append(np.ones((n, n)))
        mats.append(np.zeros((n, n)))
        # polygon with agents exactly on the slots / at the centre: symmetric ties
        ang = 2 * np.pi * np.arange(n) / n
        slots = 0.5 * np.stack([np.cos(ang), np.sin(ang)], 1)
        for agents in (slots[::-1].copy(), np.zeros((n, 2)), np.roll(slots, 1, 0)):
            mats.append(np.sqrt(((slots[None] - agents[:, None]) ** 2).sum(-1)))
    for n in (24, 32):
        for _ in range(4):
            mats.append(rng.random((n, n)))
            mats.append(rng.integers(0, 4, (n, n)).astype(np.float64))
    return mats


def make_lsa():
    mats = lsa_cases()
    out = {"scipy_version": np.array(scipy.__version__), "count": np.array(len(mats))}
    for k, m in enumerate(mats):
        r, c = linear_sum_assignment(m)
        assert (r == np.arange(m.shape[0])).all()
        out[f"cost_{k}"] = m
        out[f"col4row_{k}"] = c.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "lsa_scipy.npz"), **out)
    print("lsa_scipy.npz:", len(mats), "matrices")


TRAJ = [("navigation", 3, {}), ("navigation", 6, {"share_reward": True}),
        ("polygon", 4, {}), ("polygon", 6, {}), ("line", 5, {}),
        ("navigation", 4, {"action_mode": "continuous", "max_nbrs": 3, "own_goal_always": False})]


def make_traj(out_dir=HERE):
    for name, N, kw in TRAJ:
        cfg = worlds.make_world(name, N, dtype="f64", **kw)      # the oracle's own literal tables
        B, T = 4, 25
        rng = np.random.default_rng(zlib.crc32(f"{name}{N}".encode()))
        env = O.OracleEnv(cfg, B)
        env.reset(777)          # Philox initial states (SPEC §8)
        # squeeze the worlds so that contacts, the speed clamp and the cost fire
        ag0 = env.agent_state.copy() * 0.45
        lm0 = env.landmark_pos.copy() * 0.45
        if cfg.action_mode == "discrete":
            acts = rng.integers(0, len(cfg.discrete_u), (T, B, N)).astype(np.int32)
        else:
            acts = rng.uniform(-1, 1, (T, B, N, 2))
        pes = [py_env.PyEnv(cfg) for _ in range(B)]
        for b, p in enumerate(pes):
            p.set_state(ag0[b], lm0[b])
        rec = {k: [] for k in ("obs", "nbr_idx", "nbr_feat", "nbr_cnt", "adj", "reward", "cost",
                               "done", "assign", "agent_state")}
        for t in range(T):
            outs = [p.step(acts[t, b]) for b, p in enumerate(pes)]
            for k in rec:
                if k == "agent_state":
                    rec[k].append(np.stack([p.get_state() for p in pes]))
                else:
                    rec[k].append(np.stack([o[k] for o in outs]))
        tag = f"traj_{name}_{N}" + ("_" + "_".join(f"{a}-{b}" for a, b in sorted(kw.items())) if kw else "")
        np.savez_compressed(os.path.join(out_dir, tag + ".npz"), agent_state0=ag0, landmark_pos=lm0,
                            actions=acts, **{k: np.stack(v) for k, v in rec.items()})
        print(tag, "cost events:", int(np.stack(rec["cost"]).sum()))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--traj-only":     # tests: regenerate into a scratch directory
        make_traj(sys.argv[2])
    else:
        make_lsa()
        make_traj()
