"""A/B of the numpy-facing step (GraphVecEnv.step -> gsm_step_host) on one B200: dense D2H copy, sparse export,
sparse export with outputs switched off.   python profiles/e2e_variants.py [--steps 300]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gs_marl_b200 import scenarios  # noqa: E402
from gs_marl_b200.env_wrappers import GraphVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--envs", type=int, default=16384)
a = ap.parse_args()
cfg = scenarios.load("navigation").make_world(3, dtype="f32", episode_length=25)
vec = GraphVecEnv(cfg, a.envs, seed=1)
vec.reset()
acts = np.random.default_rng(0).integers(0, 5, (8, a.envs, 3)).astype(np.int32)
outs = [k for k in vec.buf if k != "actions"]
res = {}
for name, kw in (("dense", dict(outputs=None, sparse=False)), ("sparse_all", dict(outputs=None, sparse=True)),
                 ("sparse_no_idx", dict(outputs=[k for k in outs if k != "nbr_idx"], sparse=True)),
                 ("sparse_no_idx_assign", dict(outputs=[k for k in outs if k not in ("nbr_idx", "assign")], sparse=True)),
                 ("sparse_feat_only", dict(outputs=["nbr_feat"], sparse=True)),
                 ("sparse_no_feat", dict(outputs=[k for k in outs if k != "nbr_feat"], sparse=True)),
                 ("dense_no_feat", dict(outputs=[k for k in outs if k != "nbr_feat"], sparse=False)),
                 ("no_outputs", dict(outputs=[], sparse=True)), ("done_only", dict(outputs=["done"], sparse=True))):
    vec.set_host_outputs(**kw)
    for s in range(10):
        vec.step(acts[s % 8])
    t0 = time.perf_counter()
    for s in range(a.steps):
        vec.step(acts[s % 8])
        if s % 25 == 24:
            vec.reset()
    dt = time.perf_counter() - t0
    res[name] = {"us_per_step": round(dt / a.steps * 1e6, 1), "agent_steps_per_s": round(a.envs * 3 * a.steps / dt)}
    print(name, res[name], flush=True)
print("RESULT " + json.dumps(res))
