"""What the warm-started assignment of env_team_kernel does on a configuration (needs a -DGSM_TEAM_STATS=1 build:
make -C gs_marl_b200/csrc variant NAME=stats FLAGS=-DGSM_TEAM_STATS=1; GSM_LIB_PATH=.../variants/libstats.so)."""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gs_marl_b200 import abi, scenarios  # noqa: E402
from gs_marl_b200.environment import MultiAgentGraphConstrainEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenario", default="polygon")
ap.add_argument("--agents", type=int, default=12)
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--T", type=int, default=25)
a = ap.parse_args()
cfg = scenarios.load(a.scenario).make_world(a.agents, dtype="f32", episode_length=25)
env = MultiAgentGraphConstrainEnv(cfg, a.envs, seed=1)
env.reset()
acts = torch.randint(0, 5, (a.T, a.envs, a.agents), device="cuda", dtype=torch.int32)
ring = {k: env._alloc(k, (a.T,)) for k in env.OUTPUTS}
lib = abi.load_library()
out = (C.c_uint64 * 8)()
env.rollout(acts, out=ring, auto_reset=True)
lib.gsm_debug_team_stats(out)                      # drop the first rollout
for _ in range(3):
    env.rollout(acts, out=ring, auto_reset=True)
lib.gsm_debug_team_stats(out)
names = ("solves", "warm_attempts", "certified", "cold_solves", "free_rows", "warp_aug_rounds", "warp_cold_runs", "incomplete")
d = dict(zip(names, [int(x) for x in out]))
steps = 3 * a.T
warps = a.envs // 8
print(d)
if d["warm_attempts"]:
    print("certified / warm attempts", d["certified"] / d["warm_attempts"], " free rows per warm attempt", d["free_rows"] / d["warm_attempts"])
    print("warp-level augment rounds per warp-step", d["warp_aug_rounds"] / (warps * steps), " warp-level cold runs per warp-step", d["warp_cold_runs"] / (warps * steps))
