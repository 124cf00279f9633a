"""Helpers to load the committed golden fixtures (tests/golden/*.npz, produced by the reference itself
through tests/golden/make_golden.py)."""
import json
import os
from collections import OrderedDict

import numpy as np
import torch

from oracle import game_oracle as go

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TRAIN_CASES = ["c1_continuous", "continuous_t3", "fixed_small", "fixed_t1_noent", "adaptive_small",
               "adaptive_b1_adam", "headline_mid", "adaptive_sgd", "flipout_small", "ignore_rec_first1", "mix_prod", "ignore_code", "mix_mou", "mix_mou_ignore"]
ATTN_TRAIN_CASES = ["desc_attn_small", "desc_attn_mid"]      # -desc_attn (model.py:344-410)
ATTN_EVAL_CASES = ["eval_desc_attn"]
EVAL_CASES = ["eval_adaptive", "eval_adaptive_noprod", "eval_fixed_corrupt", "eval_continuous"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    cfg = go.GameConfig(**json.loads(str(z["cfg"])))
    return z, cfg


def params_at(z, tag, agents=go.AGENTS):
    """tag = 'P0' (initial), 'P1' (after iteration 0) ...; returns only agents stored in the fixture."""
    out = OrderedDict()
    for a in agents:
        keys = [k for k in z.files if k.startswith("%s/%s/" % (tag, a))]
        if not keys:
            continue
        out[a] = OrderedDict((k.split("/", 2)[2], torch.from_numpy(z[k].copy())) for k in keys)
    return out


def uniforms_at(z, it, cfg):
    pre = "it%d/" % it
    u_s = z[pre + "u_s"]
    steps = u_s.shape[0]
    if cfg.use_binary and (pre + "u_fz") in z.files:
        return [(z[pre + "u_z"][t], u_s[t], z[pre + "u_w"][t], z[pre + "u_fz"][t], z[pre + "u_fw"][t]) for t in range(steps)]
    if cfg.use_binary:
        return [(z[pre + "u_z"][t], u_s[t], z[pre + "u_w"][t]) for t in range(steps)]
    return [(None, u_s[t], None) for t in range(steps)]


def pad_uniforms(us, cfg, B, seed=99):
    """The reference stops drawing after an early break; kernels that run to max_exchange need draws for the
    (masked-out) remaining steps too."""
    rng = np.random.RandomState(seed)
    us = list(us)
    while len(us) < cfg.max_exchange:
        us.append((rng.rand(B, cfg.rec_w_dim), rng.rand(B, 1), rng.rand(B, cfg.rec_w_dim)))
    return us


def desc_set_at(z, it=None):
    """-desc_attn fixtures: dict(desc_set=(NW, WV) tensor, desc_set_lens=[D]) or {}."""
    pre = "" if it is None else "it%d/" % it
    if pre + "desc_set" not in z.files:
        return {}
    return dict(desc_set=torch.from_numpy(z[pre + "desc_set"].copy()),
                desc_set_lens=[int(v) for v in z[pre + "desc_set_lens"]])


def batch_at(z, it):
    pre = "it%d/" % it
    return (torch.from_numpy(z[pre + "x"].copy()), torch.from_numpy(z[pre + "desc"].copy()),
            torch.from_numpy(z[pre + "target"].copy()))
