"""Tiny training iterations (fast-path and generic shapes) for compute-sanitizer runs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodalgame_b200 import capi, engine as eng
from oracle import game_oracle as go
from tests import parity_util as pu
lib = capi.load()
for name, kw in (("fast", dict(img_h_dim=256, rec_hidden=64, sender_out_dim=32, rec_w_dim=32)),
                 # batch 64 / 128 features: tcgen05 + TMA image layer, two-level statistics, staged row tiles
                 ("fast+umma", dict(img_h_dim=256, rec_hidden=64, sender_out_dim=32, rec_w_dim=32, batch_size=64, img_feat_dim=128)),
                 ("generic+mou", dict(img_h_dim=40, rec_hidden=24, sender_out_dim=12, rec_w_dim=12, sender_mix="mou", ignore_code=True)),
                 ("generic", dict(img_h_dim=40, rec_hidden=24, sender_out_dim=12, rec_w_dim=12)),
                 ("fast+desc_attn", dict(img_h_dim=256, rec_hidden=64, sender_out_dim=32, rec_w_dim=32, desc_attn=True, desc_attn_dim=64)),
                 ("generic+desc_attn", dict(img_h_dim=40, rec_hidden=24, sender_out_dim=12, rec_w_dim=12, desc_attn=True, desc_attn_dim=10))):
    for fixed in (True, False):
        base = dict(batch_size=5, img_feat_dim=72, baseline_hid_dim=36, wv_dim=20, n_classes=9, max_exchange=3,
                    fixed_exchange=fixed, use_binary=True, entropy_s=None if fixed else 0.05, entropy_sen=0.01,
                    entropy_rec=0.02, top_k_train=2)
        base.update(kw)
        cfg = go.GameConfig(**base)
        words = pu._synth_words(cfg, 1)
        e = eng.GameEngine(pu.config_from(cfg, n_words=int(words["desc_set"].shape[0]) if words else 0), device="cuda", lib=lib)
        e.load_params(go.init_params(cfg, seed=1))
        if words:
            e.set_desc_set(**words)
        x, desc, target = go.synthetic_batch(cfg, seed=1)
        for it in range(2):
            e.train_step(x, desc, target)
        e.forward(x, desc, target, train=False)
        torch.cuda.synchronize()
        print(name, "fixed" if fixed else "adaptive", "loss_rec", e.losses()["loss_rec"])
print("SANITIZE-RUN-DONE")
