"""Shared helpers of the test-suite."""
import glob
import os

import numpy as np

from gs_marl_b200 import scenarios
from oracle import worlds as oracle_worlds

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INT_KEYS = ("nbr_idx", "nbr_cnt", "adj", "done", "assign")
REAL_KEYS = ("obs", "nbr_feat", "reward", "cost")

# name, N, kwargs — the golden trajectories (tests/golden/make_golden.py TRAJ)
GOLDEN_TRAJ = [("navigation", 3, {}), ("navigation", 6, {"share_reward": True}),
               ("polygon", 4, {}), ("polygon", 6, {}), ("line", 5, {}),
               ("navigation", 4, {"action_mode": "continuous", "max_nbrs": 3,
                                  "own_goal_always": False})]


def golden_path(name, N, kw):
    tag = f"traj_{name}_{N}" + ("_" + "_".join(f"{a}-{b}" for a, b in sorted(kw.items())) if kw else "")
    return os.path.join(GOLDEN, tag + ".npz")


def make_cfg(name, N, dtype, **kw):
    """The PRODUCT's world for (name, N, kwargs) — after checking, field by field, that it equals
    the ORACLE's own literal table for the same request (oracle/worlds.py shares no code with
    gs_marl_b200/scenarios + presets): a wrong constant, slot table or shape on either side fails
    here instead of being handed to both implementations as a common input."""
    cfg = scenarios.load(name).make_world(N, dtype=dtype, **kw)
    assert_same_world(cfg, oracle_worlds.make_world(name, N, dtype=dtype, **kw), f"{name}-{N} {kw}")
    return cfg


def oracle_world(name, N, dtype, **kw):
    return oracle_worlds.make_world(name, N, dtype=dtype, **kw)


def assert_same_world(product_cfg, oracle_world_, ctx=""):
    for f in oracle_worlds.INPUT_FIELDS + ("slot_table",):
        a, b = getattr(product_cfg, f), getattr(oracle_world_, f)
        if isinstance(a, (str, bool, int)) and not isinstance(a, float):
            assert a == b, f"{ctx}: {f}: product {a!r} != oracle table {b!r}"
        elif a is None or b is None:
            assert a is None and b is None, f"{ctx}: {f}: product {a!r} != oracle table {b!r}"
        else:
            assert np.array_equal(np.asarray(a, np.float64), np.asarray(b, np.float64)), \
                f"{ctx}: {f}: product {a!r} != oracle table {b!r}"
    n = 7
    ps, os_ = product_cfg.io_shapes(n), oracle_world_.io_shapes(n)
    assert set(ps) == set(os_), ctx
    for k in ps:
        assert np.dtype(ps[k][0]) == np.dtype(os_[k][0]) and tuple(ps[k][1]) == tuple(os_[k][1]), \
            f"{ctx}: io_shapes[{k}]: product {ps[k]} != oracle {os_[k]}"
    assert product_cfg.adj_words == oracle_world_.adj_words, ctx


def random_actions(cfg, rng, lead):
    if cfg.action_mode == "discrete":
        return rng.integers(0, len(cfg.discrete_u), tuple(lead) + (cfg.n_agents,)).astype(np.int32)
    return rng.uniform(-1, 1, tuple(lead) + (cfg.n_agents, 2)).astype(cfg.np_real)


def assert_match(got, want, *, rtol, atol, ctx="", int_exact=True, keys=None):
    """Integer/bool outputs bit-exact, real outputs within tolerance."""
    for k in (keys or INT_KEYS + REAL_KEYS):
        if k not in want or k not in got:
            continue
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, (ctx, k, g.shape, w.shape)
        if k in INT_KEYS:
            if int_exact:
                bad = np.argwhere(g != w)
                assert bad.size == 0, f"{ctx} {k}: {len(bad)} mismatches, first at {bad[0]}: {g[tuple(bad[0])]} != {w[tuple(bad[0])]}"
        else:
            np.testing.assert_allclose(g, w, rtol=rtol, atol=atol, err_msg=f"{ctx} {k}")


def near_threshold_rows(cfg, agent_state, landmark_pos, eps):
    """Bool [n_envs, N]: agents with some pair distance within eps of the sensing radius or of
    the contact distance (fp64 state).  fp32 may legitimately flip those predicates."""
    N = cfg.n_agents
    pos = np.concatenate([agent_state[..., :2], landmark_pos], axis=1).astype(np.float64)   # [B, E, 2]
    d = np.sqrt(((pos[:, None, :, :] - pos[:, :N, None, :]) ** 2).sum(-1))                    # [B, N, E]
    size = np.asarray(cfg.size, np.float64)
    dmin = size[:N, None] + size[None, :]
    near = (np.abs(d - cfg.sensing_radius) < eps) | (np.abs(d - dmin[None]) < eps)
    for i in range(N):
        near[:, i, i] = False
    return near.any(-1)


# ---- reference-recorded goldens (tools/unblock.py) ----------------------------------------------
def ref_golden_files(tmp_factory=None):
    """Trajectory files in the unblock kit's format.  GSM_REF_GOLDEN_DIR (set by `tools/unblock.py
    --pytest` once the real gsmarl/ sources are mounted) wins; without it the files are recorded
    here and now from the SYNTHETIC stand-in tree of tools/fake_gsmarl.py — an implementation of
    SPEC.md that shares no code with the product or the oracle — so the replay path itself is
    always exercised."""
    import subprocess
    import sys
    d = os.environ.get("GSM_REF_GOLDEN_DIR")
    if not d:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        base = str(tmp_factory.mktemp("refgold"))
        sys.path.insert(0, os.path.join(root, "tools"))
        import fake_gsmarl
        fake_gsmarl.write_tree(os.path.join(base, "ref"))
        subprocess.run([sys.executable, os.path.join(root, "tools", "unblock.py"), "--ref", os.path.join(base, "ref"),
                        "--out", os.path.join(base, "out"), "--configs", "nav-3,nav-6,polygon-6,line-6"],
                       check=True, capture_output=True, timeout=600)
        d = os.path.join(base, "out", "goldens")
    return sorted(glob.glob(os.path.join(d, "ref_*.npz")))


def load_ref_golden(path):
    """(recording dict, field dict of the world it was recorded in) of one ref_*.npz."""
    import json
    z = np.load(path)
    return {k: z[k] for k in z.files if k != "world_json"}, json.loads(str(z["world_json"]))


def ref_world_pair(fields, dtype="f64"):
    """The product WorldConfig and the oracle World of a recorded reference world (continuous
    controls: the recording stores the force each action produced, SPEC §2)."""
    from gs_marl_b200.config import WorldConfig
    f = dict(fields, dtype=dtype, action_mode="continuous")
    f["discrete_u"] = [tuple(u) for u in f["discrete_u"]]
    slot = oracle_worlds.spec_slot_table(f["scenario"], f["n_agents"])
    return WorldConfig(slot_table=slot, **f), oracle_worlds.World(slot_table=slot, **f)
