"""Generates tests/golden/policy_golden.npz: a small fixed input set for the declared graph actor
(SPEC.md §10) and GAE (§11) with the numpy oracle's outputs, so that a later change to
oracle/policy_oracle.py cannot silently move the target the CUDA kernels are tested against.

    python tests/golden/make_policy_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gs_marl_b200.policy import GraphAttentionActor  # noqa: E402
from oracle import policy_oracle as P  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    R, K = 96, 8
    obs = rng.standard_normal((R, 6)).astype(np.float32)
    cnt = rng.integers(0, K + 1, R).astype(np.int32)
    feat = rng.standard_normal((R, K, 6)).astype(np.float32)
    feat[np.arange(K)[None, :] >= cnt[:, None]] = 0
    actor = GraphAttentionActor(5, seed=2026)
    w = P.weights_from_state_dict(actor.state_dict())
    a, lp, z, margin = P.act(w, obs, feat, cnt, seed=99, step=7, row_offset=12345)
    v = P.values(w, obs, feat, cnt)
    T = 6
    rw, cs = rng.standard_normal((T, R)), rng.random((T, R))
    vals = rng.standard_normal((T + 1, R, 2))
    done = (rng.random((T, R)) < 0.15).astype(np.uint8)
    ret, adv = P.gae(rw, cs, vals, done, 0.99, 0.95)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "policy_golden.npz"),
                        obs=obs, cnt=cnt, feat=feat, actions=a, logp=lp, logits=z, margin=margin, values=v,
                        gumbel=P.gumbel(R, 5, 99, 7, 12345), gae_reward=rw, gae_cost=cs, gae_values=vals,
                        gae_done=done, gae_returns=ret, gae_adv=adv,
                        **{"w_" + k: np.asarray(val) for k, val in w.items()})


if __name__ == "__main__":
    main()
