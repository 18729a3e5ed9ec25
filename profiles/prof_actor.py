"""ncu target: the actor kernel on one observation slot of the bench workload (nav-3 x 16384)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gs_marl_b200 import scenarios
from gs_marl_b200.environment import MultiAgentGraphConstrainEnv
from gs_marl_b200.policy import GraphAttentionActor

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
cfg = scenarios.load("navigation").make_world(N, dtype="f32", episode_length=25)
env = MultiAgentGraphConstrainEnv(cfg, envs, device=0, seed=1)
obs, graph = env.reset()
actor = GraphAttentionActor(5)
for s in range(10):
    a, lp = actor.act(obs, graph, seed=1, step=s)
    obs, graph, *_ = env.step(a)
torch.cuda.synchronize()
out = (torch.empty_like(graph["nbr_cnt"]), torch.empty_like(obs[..., 0]))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for s in range(50):
        actor.act(obs, graph, seed=1, step=s, out=out)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(4):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("actor_us", e0.elapsed_time(e1) * 1e3 / 200, "mean rows", graph["nbr_cnt"].float().mean().item())
