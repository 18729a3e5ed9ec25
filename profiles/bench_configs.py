"""Device-resident step time of every BASELINE.json configuration on one GPU.

    python profiles/bench_configs.py [--dtype f32] [--json out.json] [--only substr]

Prints two lines per configuration — one handle on one stream, and `--streams` sub-shards on
concurrent streams (StreamShardedEnv) — each with step time, agent-steps/s, algorithmic GB/s and
the fraction of the measured HBM peak.  Uses gsm_rollout with in-kernel auto-reset (fused launch
where a specialised kernel exists, CUDA graph of generic steps otherwise) into a T-slot buffer.
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gs_marl_b200 import scenarios  # noqa: E402
from gs_marl_b200.environment import MultiAgentGraphConstrainEnv, StreamShardedEnv  # noqa: E402

CONFIGS = [  # (label, scenario, N, n_envs, kwargs)
    ("cfg1 nav-3 x16384", "navigation", 3, 16384, {}),
    ("cfg2 nav-24 x4096", "navigation", 24, 4096, {"max_nbrs": 32}),
    ("cfg2 nav-48 x4096", "navigation", 48, 4096, {"max_nbrs": 32}),
    ("cfg2 nav-96 x4096", "navigation", 96, 4096, {"max_nbrs": 32}),
    ("cfg3 polygon-6 x16384", "polygon", 6, 16384, {}),
    ("cfg3 polygon-12 x16384", "polygon", 12, 16384, {}),
    ("cfg3 line-6 x16384", "line", 6, 16384, {}),
    ("cfg3 line-12 x16384", "line", 12, 16384, {}),
    ("cfg4 nav-12 x8192 (65536/8 GPUs)", "navigation", 12, 8192, {}),
    ("cfg4 nav-12 x65536", "navigation", 12, 65536, {}),
    ("nav-6 x16384", "navigation", 6, 16384, {}),
    ("nav-4 x16384", "navigation", 4, 16384, {}),
    ("nav-5 x16384", "navigation", 5, 16384, {}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    ap.add_argument("--streams", type=int, default=4, help="sub-shards on concurrent streams (second column)")
    args = ap.parse_args()
    peak = 6546.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    out = []
    for label, scn, N, n_envs, kw in CONFIGS:
        if args.only and args.only not in label:
            continue
        cfg = scenarios.load(scn).make_world(N, dtype=args.dtype, **kw)
        bpas = cfg.bytes_per_agent_step()
        T = 8 if bpas * N * n_envs * 25 > 8e9 else 25
        acts = torch.randint(0, 5, (T, n_envs, N), device="cuda", dtype=torch.int32)
        ring = None
        rec = {"config": label, "dtype": args.dtype, "bytes_per_agent_step": bpas, "steps_per_rollout": T}
        # both columns run the steady state of the bench: auto-reset on, episode_length of the preset
        for col, S in (("one_stream", 1), (f"streams_{args.streams}", args.streams)):
            env = (MultiAgentGraphConstrainEnv(cfg, n_envs, seed=3) if S == 1 else
                   StreamShardedEnv(cfg, n_envs, n_streams=S, seed=3))
            env.reset()
            if ring is None:
                ring = {k: env._alloc(k, (T,)) for k in env.OUTPUTS}
            if S == 1:
                io = env._make_io(ring, acts)
                env._check(env.lib.gsm_set_auto_reset(env._h, 1))
                st = env._stream()

                def launch(env=env, io=io, st=st):
                    env._check(env.lib.gsm_rollout(env._h, T, C.byref(io), st))
                join = lambda: None
            else:
                launch, join = env.rollout_plan(acts, ring, auto_reset=True), env.join
            for _ in range(4):
                launch()
            join()
            torch.cuda.synchronize()
            reps = 8
            l0 = env.kernel_launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if S > 1:
                for s_ in env.streams:
                    s_.wait_event(e0)
            for _ in range(reps):
                launch()
            join()
            e1.record()
            torch.cuda.synchronize()
            step_us = e0.elapsed_time(e1) * 1e3 / (reps * T)
            gbs = bpas * N * n_envs / (step_us * 1e-6) / 1e9
            rec[col] = {"step_us": step_us, "agent_steps_per_s": N * n_envs / (step_us * 1e-6),
                        "achieved_gbs": gbs, "frac_of_measured_hbm": gbs / peak,
                        "launches_per_rollout": (env.kernel_launches - l0) // reps}
            print(f"{label:36s} {col:11s} {step_us:9.2f} us/step  {rec[col]['agent_steps_per_s']:.3e} agent-steps/s  "
                  f"{gbs:7.1f} GB/s  {gbs / peak:6.1%}  launches/rollout={rec[col]['launches_per_rollout']}", flush=True)
            env.close()
        out.append(rec)
        del ring, acts
        torch.cuda.empty_cache()
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
