"""Intra-GPU env sub-shards on concurrent streams: does overlapping the tail of one shard's fused
rollout with the body of another's raise the aggregate step rate at the bench batch?

One B200, `total` envs split into S contiguous shards (same global env indices as one handle:
env_offset = shard start, so the worlds are identical), each shard on its own CUDA stream, R
rounds of T-step fused rollouts per shard enqueued round-robin; timed with CUDA events on the
default stream (fork: every stream waits for ev0; join: the default stream waits for every stream).

    python profiles/sweep_streams.py [scenario N total_envs [T]]
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gs_marl_b200 import scenarios  # noqa: E402
from gs_marl_b200.environment import MultiAgentGraphConstrainEnv  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def run(scn, N, total, T, S, auto, rounds=40, episode_length=100):
    dev = torch.device("cuda", 0)
    cfg = scenarios.load(scn).make_world(N, dtype="f32", episode_length=episode_length)
    bounds = [(total * s) // S for s in range(S + 1)]
    shards = []
    for s in range(S):
        n = bounds[s + 1] - bounds[s]
        env = MultiAgentGraphConstrainEnv(cfg, n, device=0, env_offset=bounds[s], seed=1)
        env.reset()
        acts = torch.randint(0, 5, (T, n, N), device=dev, dtype=torch.int32)
        ring = {k: env._alloc(k, (T,)) for k in env.OUTPUTS}
        io = env._make_io(ring, acts)
        env._check(env.lib.gsm_set_auto_reset(env._h, int(auto)))
        st = torch.cuda.Stream(device=dev)
        shards.append((env, io, st, ring, acts))
    torch.cuda.synchronize()
    main = torch.cuda.current_stream(dev)

    def go(r):
        for _ in range(r):
            for env, io, st, _, _ in shards:
                env._check(env.lib.gsm_rollout(env._h, T, C.byref(io), C.c_void_p(st.cuda_stream)))

    go(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _, _, st, _, _ in shards:
        st.wait_event(e0)
    go(rounds)
    for _, _, st, _, _ in shards:
        main.wait_stream(st)
    e1.record(main)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (rounds * T)
    gbs = cfg.bytes_per_agent_step() * total * N / (us * 1e-6) / 1e9
    for env, *_ in shards:
        env.close()
    return {"scenario": scn, "N": N, "envs": total, "T": T, "streams": S, "auto_reset": bool(auto),
            "step_us": round(us, 3), "GBs": round(gbs, 1), "frac": round(gbs / PEAK, 4),
            "agent_steps_per_s": total * N / (us * 1e-6)}


if __name__ == "__main__":
    scn = sys.argv[1] if len(sys.argv) > 1 else "navigation"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    total = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    for auto in (1, 0):
        for S in (1, 2, 3, 4, 6, 8):
            print(json.dumps(run(scn, N, total, T, S, auto)), flush=True)
