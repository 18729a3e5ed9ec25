"""BASELINE.json configs[4]: env-sharded closed-loop rollout, 65536 envs x 12 agents over the
GPUs of one box (STRONG scaling: the global env count is fixed), a random-init graph actor in the
loop on every GPU, nothing leaves the device, no collective on the step path (one final stats
all-reduce).  Launch like bench.py:

    python profiles/bench_config5.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 profiles/bench_config5.py

Two arms per run: `torch_policy` = the same actor architecture in plain PyTorch (library kernels)
between gsm_step launches (rollout.collect), and `fused_actor` = this library's actor kernel via
gsm_collect, CUDA-graphed (rollout.collect_fused).  Time = CUDA events, max over ranks.
The model is the declared one of SPEC.md (GS-MARL's env and actor are withheld): UNVERIFIED presets.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gs_marl_b200 import scenarios  # noqa: E402
from gs_marl_b200.env_wrappers import ShardedStats, shard_bounds  # noqa: E402
from gs_marl_b200.environment import MultiAgentGraphConstrainEnv, StreamShardedEnv  # noqa: E402
from gs_marl_b200.policy import GraphAttentionActor  # noqa: E402
from gs_marl_b200.rollout import GraphRolloutBuffer, collect, collect_fused  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="GLOBAL env count")
    ap.add_argument("--agents", type=int, default=12)
    ap.add_argument("--rollouts", type=int, default=8)
    ap.add_argument("--T", type=int, default=25)
    ap.add_argument("--streams", type=int, default=1, help="env sub-shards per GPU for the fused arm")
    ap.add_argument("--arms", default="fused_actor,torch_policy")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)           # NCCL banner must not reach stdout
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    lo, hi = shard_bounds(args.envs, world, rank)
    n_envs, N, T = hi - lo, args.agents, args.T
    cfg = scenarios.load("navigation").make_world(N, dtype="f32", episode_length=T)
    actor = GraphAttentionActor(len(cfg.discrete_u), seed=0)
    torch_actor = GraphAttentionActor(len(cfg.discrete_u), seed=0).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(arm):
        if arm == "fused_actor" and args.streams > 1:
            env = StreamShardedEnv(cfg, n_envs, n_streams=args.streams, device=local, env_offset=lo, seed=11)
        else:
            env = MultiAgentGraphConstrainEnv(cfg, n_envs, device=local, env_offset=lo, seed=11)
        buf = GraphRolloutBuffer(env, T)
        buf.reset_env()
        if arm == "fused_actor":
            g = collect_fused(env, actor, buf, seed=5, graph=True, with_values=False)   # same work as the torch arm
            once = g.replay
        else:
            @torch.no_grad()
            def policy(obs, graph):
                z = torch_actor.logits_autograd(obs, graph)
                gmb = -torch.log(-torch.log(torch.rand_like(z).clamp_min(1e-9)))
                return (z + gmb).argmax(-1).to(torch.int32)

            def once():
                collect(env, policy, buf)
        for _ in range(2):
            once(); buf.reset_env()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.rollouts):
            once()
            buf.reset_env()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        stats = ShardedStats()
        stats.add(n_envs * T, N, float(buf["reward"].sum()), float(buf["cost"].sum()), float(buf["done"].sum()))
        tot = stats.all_reduce(device=dev)
        launches = env.kernel_launches
        env.close()
        steps = args.rollouts * T
        return {"value": args.envs * N * steps / (ms.item() * 1e-3), "unit": "agent-steps/s",
                "ms_per_step": ms.item() / steps, "steps": steps, "last_rollout_stats": tot,
                "gsm_launches_rank0": launches}

    res = {arm: run(arm) for arm in args.arms.split(",")}
    res["fused_streams"] = args.streams
    if rank == 0:
        print(json.dumps({
            "config": f"BASELINE configs[4]: navigation, {N} agents, {args.envs} envs sharded over {world} GPU(s) "
                      f"({n_envs} per GPU), closed loop with a random-init graph actor, fp32",
            "n_gpus": world, "scaling": "strong", "envs_per_gpu": n_envs, "rollout_T": T,
            "spec_status": "declared model (SPEC.md), UNVERIFIED presets; reference env and actor withheld",
            **res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
