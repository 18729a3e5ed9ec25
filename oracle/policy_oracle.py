"""TEST INFRASTRUCTURE — numpy restatement of the declared graph actor (SPEC.md §10) that
gs_marl_b200/csrc/gsm_policy.cu implements.  Only tests/, smoke() and
bench.py's checker legs may import this.

PARITY UNPINNED: the reference's GNN actor (gsmarl/algorithms/*, torch-geometric per
requirements.txt:119) is withheld, so this restates SPEC.md §10, not GS-MARL.  What IS pinned:
Philox4x32-10 (to the C oracle's, which is pinned to the Random123 known-answer vectors) and
the forward pass to an independent plain-torch fp32 formulation (tests/test_policy.py).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over equal-shaped uint32 counter arrays; scalar key.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = h1 ^ c1 ^ np.uint64(k0), l1, h0 ^ c3 ^ np.uint64(k1), l0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def weights_from_state_dict(sd, dtype=np.float32):
    g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k], dtype=dtype)
    return {"ego_w": g("ego.weight"), "ego_b": g("ego.bias"), "nbr_w": g("nbr.weight"), "nbr_b": g("nbr.bias"),
            "att_w": g("att.weight").reshape(-1), "att_b": g("att.bias").reshape(()),
            "head_w": g("head.weight"), "head_b": g("head.bias"),
            "value_w": g("value.weight"), "value_b": g("value.bias")}


def logits(w, obs, nbr_feat, nbr_cnt, dtype=np.float64):
    """SPEC.md §10 forward.  obs [R,6], nbr_feat [R,K,6], nbr_cnt [R] -> [R, n_actions]."""
    return forward(w, obs, nbr_feat, nbr_cnt, dtype)[0]


def values(w, obs, nbr_feat, nbr_cnt, dtype=np.float64):
    """The two critics (reward value, cost value) on the same embedding -> [R, 2]."""
    return forward(w, obs, nbr_feat, nbr_cnt, dtype)[1]


def forward(w, obs, nbr_feat, nbr_cnt, dtype=np.float64):
    w = {k: np.asarray(v, dtype=dtype) for k, v in w.items()}
    obs, feat = np.asarray(obs, dtype=dtype), np.asarray(nbr_feat, dtype=dtype)
    R, K = feat.shape[0], feat.shape[1]
    Hh = w["ego_b"].shape[0]
    cnt = np.clip(np.asarray(nbr_cnt), 0, K)
    e = np.maximum(obs @ w["ego_w"].T + w["ego_b"], 0)
    m = np.maximum(feat @ w["nbr_w"].T + w["nbr_b"], 0)                 # [R,K,H]
    sc = m @ w["att_w"] + w["att_b"]                                    # [R,K]
    valid = np.arange(K)[None, :] < cnt[:, None]
    sc = np.where(valid, sc, -np.inf)
    mx = np.where(cnt > 0, sc.max(1, initial=-np.inf), 0.0)
    p = np.where(valid, np.exp(sc - mx[:, None]), 0.0)
    s = p.sum(1)
    a = p / np.where(s > 0, s, 1.0)[:, None]
    agg = (a[..., None] * m).sum(1)                                     # zero when cnt == 0
    emb = np.concatenate([e, agg], 1)
    return emb @ w["head_w"].T + w["head_b"], emb @ w["value_w"].T + w["value_b"]


def gumbel(n_rows, n_actions, seed, step, row_offset=0):
    """fp32 Gumbel noise the kernel adds: Philox block b covers actions 4b..4b+3; counter
    (row lo, row hi, step, 0x80000000 | b), key (seed lo, seed hi); u = (x >> 9) * 2^-23 + 2^-24 (exact in fp32, inside (0, 1))."""
    g = np.uint64(row_offset) + np.arange(n_rows, dtype=np.uint64)
    out = np.empty((n_rows, n_actions), dtype=np.float32)
    for b in range((n_actions + 3) // 4):
        r = philox4x32_10(g & MASK, g >> np.uint64(32), np.full(n_rows, step & 0xFFFFFFFF, np.uint64),
                          np.full(n_rows, 0x80000000 | b, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for q in range(4):
            a = 4 * b + q
            if a < n_actions:
                u = (r[q] >> np.uint32(9)).astype(np.float32) * np.float32(2.0 ** -23) + np.float32(2.0 ** -24)
                out[:, a] = -np.log(-np.log(u))
    return out


def act(w, obs, nbr_feat, nbr_cnt, seed=0, step=0, row_offset=0, greedy=False, dtype=np.float64):
    """-> actions int32 [R], logp [R], logits [R,A], margin [R] (top-2 gap of the perturbed
    logits: rows with a tiny margin may legitimately flip under fp32 rounding)."""
    z = logits(w, obs, nbr_feat, nbr_cnt, dtype)
    v = z if greedy else z + gumbel(z.shape[0], z.shape[1], seed, step, row_offset).astype(z.dtype)
    a = v.argmax(1).astype(np.int32)
    srt = np.sort(v, 1)
    margin = srt[:, -1] - srt[:, -2]
    zm = z.max(1, keepdims=True)
    lse = zm[:, 0] + np.log(np.exp(z - zm).sum(1))
    return a, z[np.arange(z.shape[0]), a] - lse, z, margin


def gae(reward, cost, values, done, gamma, lam, dtype=np.float64):
    """SPEC.md §11.  reward, cost, done [T, R]; values [T+1, R, 2] -> returns, advantages [T, R, 2]."""
    rw = np.stack([np.asarray(reward, dtype), np.asarray(cost, dtype)], -1)
    v = np.asarray(values, dtype)
    mask = (1.0 - (np.asarray(done) != 0).astype(dtype))[..., None]
    T = rw.shape[0]
    ret, adv = np.zeros_like(rw), np.zeros_like(rw)
    g = np.zeros_like(rw[0])
    for t in range(T - 1, -1, -1):
        delta = rw[t] + gamma * v[t + 1] * mask[t] - v[t]
        g = delta + gamma * lam * mask[t] * g
        adv[t], ret[t] = g, g + v[t]
    return ret, adv
