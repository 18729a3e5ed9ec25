"""A/B sweep of the nav-3 hot kernel on one B200: every variant (an alternative build of the library
and/or GSM_* environment switches) runs in its own process; per variant the bench's timed region
(CUDA-graph replay, bench_util.RolloutRegion) with in-kernel auto-reset (MODE 2) and without
(MODE 0), on 4 sub-shard streams and as one launch on one stream.

    python profiles/sweep_nav3.py --variants base,P2:GSM_SPEC_P=2,x:lib=gs_marl_b200/csrc/variants/libx.so
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(args):
    sys.path.insert(0, ROOT)
    import torch
    from bench_util import time_rollouts
    from gs_marl_b200 import scenarios
    from gs_marl_b200.environment import MultiAgentGraphConstrainEnv, StreamShardedEnv
    N, n_envs, T = args.agents, args.envs, args.T
    cfg = scenarios.load(args.scenario).make_world(N, dtype="f32", episode_length=25)
    dev = torch.device("cuda", 0)
    acts = torch.randint(0, 5, (T, n_envs, N), device=dev, dtype=torch.int32)
    one = MultiAgentGraphConstrainEnv(cfg, n_envs, seed=1)
    one.reset()
    ring = {k: one._alloc(k, (T,)) for k in one.OUTPUTS}
    sh = StreamShardedEnv(cfg, n_envs, n_streams=args.streams, seed=1)
    sh.reset()
    peak = 6546.2
    out = {}
    for name, env in (("s4", sh), ("s1", one)):
        for mode, auto in (("auto", True), ("plain", False)):
            r = time_rollouts(env, cfg, acts, ring, args.K, T, auto, args.ms)
            out[f"{name}_{mode}_us"] = round(r["step_us"], 3)
            out[f"{name}_{mode}_frac"] = round(r["achieved"] / peak, 4)
    # checksum of the last buffer so that variants can be compared for equality of results
    torch.cuda.synchronize()
    out["checksum"] = float(ring["nbr_feat"].double().sum().item() + ring["reward"].double().sum().item()
                            + ring["nbr_idx"].double().sum().item() + ring["obs"].double().sum().item())
    print("RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="base")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--scenario", default="navigation")
    ap.add_argument("--agents", type=int, default=3)
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--T", type=int, default=25)
    ap.add_argument("--K", type=int, default=25)
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--ms", type=float, default=150.0)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    if args.child:
        return child(args)
    res = {}
    for v in args.variants.split(","):
        name, *kvs = v.split(":")
        env, streams = dict(os.environ), args.streams
        for kv in kvs:
            k, val = kv.split("=", 1)
            if k == "lib":
                env["GSM_LIB_PATH"] = os.path.join(ROOT, val)
            elif k == "streams":
                streams = int(val)
            else:
                env[k] = val
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--scenario", args.scenario, "--agents",
               str(args.agents), "--envs", str(args.envs), "--T", str(args.T), "--K", str(args.K), "--streams",
               str(streams), "--ms", str(args.ms)]
        p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
        res[name] = json.loads(line[0][7:]) if line else {"error": (p.stderr or p.stdout)[-400:]}
        print(f"{name:14s} {json.dumps(res[name])}", flush=True)
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
