"""A few fused auto-reset rollouts of one configuration, one launch each, for ncu:
    ncu --set full --import-source on --clock-control none -k regex:env_ -s 2 -c 1 -o out \
        python profiles/prof_env.py [--scenario navigation --agents 3 --envs 16384 --T 25]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gs_marl_b200 import scenarios  # noqa: E402
from gs_marl_b200.environment import MultiAgentGraphConstrainEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenario", default="navigation")
ap.add_argument("--agents", type=int, default=3)
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--T", type=int, default=25)
ap.add_argument("--rollouts", type=int, default=4)
ap.add_argument("--max-nbrs", type=int, default=None)
a = ap.parse_args()
kw = {} if a.max_nbrs is None else {"max_nbrs": a.max_nbrs}
cfg = scenarios.load(a.scenario).make_world(a.agents, dtype="f32", episode_length=25, **kw)
env = MultiAgentGraphConstrainEnv(cfg, a.envs, seed=1)
env.reset()
acts = torch.randint(0, 5, (a.T, a.envs, a.agents), device="cuda", dtype=torch.int32)
ring = {k: env._alloc(k, (a.T,)) for k in env.OUTPUTS}
for _ in range(a.rollouts):
    env.rollout(acts, out=ring, auto_reset=True)
torch.cuda.synchronize()
print("ok", float(ring["reward"].sum()))
