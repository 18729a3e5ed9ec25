"""nav-3 at GPU-saturating env counts for every compiled lanes-per-agent variant."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from gs_marl_b200 import scenarios
from gs_marl_b200.environment import MultiAgentGraphConstrainEnv
n = int(sys.argv[1]); T = 8
cfg = scenarios.load("navigation").make_world(3, dtype="f32")
env = MultiAgentGraphConstrainEnv(cfg, n, seed=1); env.reset()
acts = torch.randint(0, 5, (T, n, 3), device="cuda", dtype=torch.int32)
ring = {k: env._alloc(k, (T,)) for k in env.OUTPUTS}
env.rollout(acts, out=ring); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): env.rollout(acts, out=ring)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * T)
gbs = cfg.bytes_per_agent_step() * 3 * n / (us * 1e-6) / 1e9
print("%%d envs  %%8.2f us/step  %%7.1f GB/s  %%5.1f%%%% of 6546" %% (n, us, gbs, 100 * gbs / 6546.2))
''' % ROOT
variants = [("spec P=%d" % P, {"GSM_SPEC_P": str(P)}) for P in (8, 4, 2, 1)] + [("lane", {"GSM_LANE_MIN_N": "3"})]
if len(sys.argv) > 1:
    variants = [v for v in variants if v[0].split()[0] in sys.argv[1:] or v[0] in sys.argv[1:]]
for n in (16384, 65536, 262144, 1048576):
    for name, ev in variants:
        env = dict(os.environ, **ev)
        out = subprocess.run([sys.executable, "-c", code, str(n)], env=env, capture_output=True, text=True)
        print(f"{name:9s}", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
