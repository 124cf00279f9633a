"""Generate golden vectors by running the UNMODIFIED reference (`/root/reference/model.py`) under
`oracle/ref_shim.py`.  Build-container only (needs /root/reference); the `.npz` files it writes are
committed so that tests on the GPU box never touch the reference.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

For every case: the reference's own `Sender`/`Receiver`/`Baseline` modules are loaded with the stored
parameters, its own `exchange()` is called, and (train cases) its own update block
(`model.py:1243-1330`, read from the reference file at run time) is executed with torch.optim — so every
stored output, loss and post-step parameter is produced by reference code, not by our oracle.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.optim as optim

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import game_oracle as go  # noqa: E402
from oracle import ref_shim as rs  # noqa: E402

CASES = {
    # C1 of BASELINE.json at full size: fixed 1-step, B=8, 5 classes, F=256, continuous messages.
    "c1_continuous": dict(cfg=dict(batch_size=8, img_feat_dim=256, img_h_dim=256, baseline_hid_dim=500,
                                   sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100, n_classes=5,
                                   max_exchange=1, fixed_exchange=True, use_binary=False, top_k_train=2),
                          iters=2, seed=11, store_params=("receiver",)),
    # continuous, several steps (raw scores travel between agents)
    "continuous_t3": dict(cfg=dict(batch_size=5, img_feat_dim=24, img_h_dim=16, baseline_hid_dim=12,
                                   sender_out_dim=8, rec_hidden=12, rec_w_dim=8, wv_dim=20, n_classes=4,
                                   max_exchange=3, fixed_exchange=True, use_binary=False, top_k_train=2),
                          iters=2, seed=12),
    # fixed-length binary exchange, small dims, 3 iterations
    "fixed_small": dict(cfg=dict(batch_size=6, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                 sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                 max_exchange=4, fixed_exchange=True, use_binary=True, entropy_sen=0.01,
                                 entropy_rec=0.01, top_k_train=3),
                        iters=3, seed=13),
    # fixed, no entropy regulariser (entropy_* = None is the flag default, model.py:1730-1732), T=1
    "fixed_t1_noent": dict(cfg=dict(batch_size=4, img_feat_dim=16, img_h_dim=8, baseline_hid_dim=8,
                                    sender_out_dim=8, rec_hidden=8, rec_w_dim=8, wv_dim=12, n_classes=3,
                                    max_exchange=1, fixed_exchange=True, use_binary=True, top_k_train=1),
                           iters=2, seed=14),
    # adaptive exchange (STOP bit), small dims, 3 iterations
    "adaptive_small": dict(cfg=dict(batch_size=9, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                    sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                    max_exchange=5, fixed_exchange=False, use_binary=True, entropy_sen=0.01,
                                    entropy_rec=0.01, entropy_s=0.08, top_k_train=3),
                           iters=3, seed=15),
    # adaptive, batch of one (std() branch `logs.size(0) > 1` not taken), Adam
    "adaptive_b1_adam": dict(cfg=dict(batch_size=1, img_feat_dim=16, img_h_dim=8, baseline_hid_dim=8,
                                      sender_out_dim=8, rec_hidden=8, rec_w_dim=8, wv_dim=12, n_classes=3,
                                      max_exchange=4, fixed_exchange=False, use_binary=True, entropy_s=0.05,
                                      top_k_train=1, optim_type="Adam", learning_rate=1e-3),
                             iters=3, seed=16),
    # headline structure (B=64, 30 classes, T=10, M=32, Hr=64, Hi=256) with F and Hb reduced to keep the file small
    "headline_mid": dict(cfg=dict(batch_size=64, img_feat_dim=128, img_h_dim=256, baseline_hid_dim=64,
                                  sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100, n_classes=30,
                                  max_exchange=10, fixed_exchange=True, use_binary=True, entropy_sen=0.01,
                                  entropy_rec=0.01, top_k_train=6),
                         iters=1, seed=17),
    # SGD, adaptive, larger batch
    "adaptive_sgd": dict(cfg=dict(batch_size=32, img_feat_dim=32, img_h_dim=16, baseline_hid_dim=16,
                                  sender_out_dim=16, rec_hidden=16, rec_w_dim=16, wv_dim=20, n_classes=10,
                                  max_exchange=6, fixed_exchange=False, use_binary=True, entropy_sen=0.01,
                                  entropy_rec=0.01, entropy_s=0.08, top_k_train=3, optim_type="SGD",
                                  learning_rate=1e-2),
                         iters=2, seed=18),
    # -ignore_receiver (receiver messages zeroed, model.py:470-472) with a non-zero first receiver message (-first_rec 1)
    "ignore_rec_first1": dict(cfg=dict(batch_size=6, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                       sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                       max_exchange=3, fixed_exchange=True, use_binary=True, entropy_sen=0.01,
                                       entropy_rec=0.02, top_k_train=2, ignore_receiver=True, first_rec=1.0),
                              iters=2, seed=20),
    # -sender_mix prod: tanh(h_x * h_w) (model.py:217-218)
    "mix_prod": dict(cfg=dict(batch_size=6, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                              sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                              max_exchange=4, fixed_exchange=False, use_binary=True, entropy_sen=0.01,
                              entropy_rec=0.02, entropy_s=0.05, top_k_train=2, sender_mix="prod"),
                     iters=2, seed=25),
    # -sender_mix mou: binary_layer reads tanh([h_x ; h_w ; h_x - h_w ; h_x * h_w]) (4 x h_dim, model.py:71-76, 219-221)
    "mix_mou": dict(cfg=dict(batch_size=6, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                             sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                             max_exchange=4, fixed_exchange=False, use_binary=True, entropy_sen=0.01,
                             entropy_rec=0.02, entropy_s=0.05, top_k_train=2, sender_mix="mou"),
                    iters=2, seed=31),
    # -sender_mix mou -ignore_code: after step 0 the code term is code_layer(sigmoid(code_bias_mou)), a second learned code
    # (model.py:73-74, 201-205, 211-213)
    "mix_mou_ignore": dict(cfg=dict(batch_size=6, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                    sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                    max_exchange=3, fixed_exchange=True, use_binary=True, entropy_sen=0.01,
                                    entropy_rec=0.02, top_k_train=2, sender_mix="mou", ignore_code=True),
                           iters=2, seed=32),
    # -ignore_code: the sender never sees the receiver's message (model.py:208-210); code_layer / code_bias get no gradient
    "ignore_code": dict(cfg=dict(batch_size=6, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                 sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                 max_exchange=3, fixed_exchange=True, use_binary=True, entropy_sen=0.01,
                                 entropy_rec=0.02, top_k_train=2, ignore_code=True),
                        iters=2, seed=26),
    # flipout noise on both messages (model.py:233-234,467-468,554-568): two extra uniform draws per step
    "flipout_small": dict(cfg=dict(batch_size=8, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                   sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                   max_exchange=4, fixed_exchange=False, use_binary=True, entropy_sen=0.01,
                                   entropy_rec=0.02, entropy_s=0.05, top_k_train=2, flipout_sen=0.15, flipout_rec=0.1),
                          iters=2, seed=19),
    # -desc_attn (model.py:344-410): ragged word-level attention, adaptive length so that prediction steps differ
    "desc_attn_small": dict(cfg=dict(batch_size=7, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                     sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=6,
                                     max_exchange=4, fixed_exchange=False, use_binary=True, entropy_sen=0.01,
                                     entropy_rec=0.02, entropy_s=0.05, top_k_train=2, desc_attn=True, desc_attn_dim=10),
                            iters=3, seed=27),
    # -desc_attn at the headline agent shapes (fast-path dimensions), fixed length, 30 classes x 3..14 words
    "desc_attn_mid": dict(cfg=dict(batch_size=16, img_feat_dim=64, img_h_dim=256, baseline_hid_dim=32,
                                   sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100, n_classes=30,
                                   max_exchange=5, fixed_exchange=True, use_binary=True, entropy_sen=0.01,
                                   entropy_rec=0.01, top_k_train=6, desc_attn=True, desc_attn_dim=64),
                          iters=1, seed=28, words=(3, 14)),
}

EVAL_CASES = {
    # eval mode: round(), running product of stop probabilities, early break, message corruption
    "eval_adaptive": dict(cfg=dict(batch_size=10, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                   sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                   max_exchange=6, fixed_exchange=False, use_binary=True),
                          seed=21, corrupt_region=None, s_bias=0.4),
    "eval_adaptive_noprod": dict(cfg=dict(batch_size=10, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                          sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                          max_exchange=6, fixed_exchange=False, use_binary=True, s_prob_prod=False),
                                 seed=22, corrupt_region=None, s_bias=0.2),
    "eval_fixed_corrupt": dict(cfg=dict(batch_size=7, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                        sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                        max_exchange=3, fixed_exchange=True, use_binary=True),
                               seed=23, corrupt_region="0:3,5", s_bias=0.0),
    "eval_continuous": dict(cfg=dict(batch_size=7, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                     sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=7,
                                     max_exchange=3, fixed_exchange=True, use_binary=False),
                            seed=24, corrupt_region=None, s_bias=0.0),
    "eval_desc_attn": dict(cfg=dict(batch_size=9, img_feat_dim=40, img_h_dim=24, baseline_hid_dim=20,
                                    sender_out_dim=12, rec_hidden=16, rec_w_dim=12, wv_dim=20, n_classes=6,
                                    max_exchange=5, fixed_exchange=False, use_binary=True, desc_attn=True,
                                    desc_attn_dim=10),
                           seed=29, corrupt_region=None, s_bias=0.3),
}


def _ref_modules(model, cfg, params):
    F = model.FLAGS
    rs.set_flags(model, max_exchange=cfg.max_exchange, use_binary=cfg.use_binary,
                 fixed_exchange=cfg.fixed_exchange, entropy_s=cfg.entropy_s, entropy_sen=cfg.entropy_sen,
                 entropy_rec=cfg.entropy_rec, batch_size=cfg.batch_size, top_k_train=cfg.top_k_train,
                 first_rec=cfg.first_rec, s_prob_prod=cfg.s_prob_prod, debug=False, sender_mix=cfg.sender_mix,
                 ignore_code=cfg.ignore_code, desc_attn=cfg.desc_attn, desc_attn_dim=cfg.desc_attn_dim, ignore_receiver=cfg.ignore_receiver, flipout_sen=cfg.flipout_sen,
                 flipout_rec=cfg.flipout_rec, flipout_dev=False, cuda=False, rec_w_dim=cfg.rec_w_dim, sender_out_dim=cfg.sender_out_dim)
    sender = model.Sender("avgpool_512", cfg.img_feat_dim, cfg.img_h_dim, cfg.rec_w_dim, cfg.sender_out_dim,
                          cfg.use_binary, False, 0, False, 0)
    receiver = model.Receiver(cfg.sender_out_dim, cfg.wv_dim, cfg.rec_hidden, 1, cfg.rec_w_dim, 1, cfg.use_binary)
    bsen = model.Baseline(cfg.baseline_hid_dim, cfg.img_h_dim, cfg.rec_w_dim, 0)
    brec = model.Baseline(cfg.baseline_hid_dim, 0, cfg.rec_w_dim, cfg.rec_hidden)
    mods = dict(receiver=receiver, sender=sender, baseline_rec=brec, baseline_sen=bsen)
    for a, m in mods.items():
        m.load_state_dict(params[a])
    return mods


def _opt(mods, cfg):
    cls = dict(RMSprop=optim.RMSprop, Adam=optim.Adam, SGD=optim.SGD)[cfg.optim_type]
    return {a: cls(m.parameters(), lr=cfg.learning_rate) for a, m in mods.items()}


def _stack(lst):
    return np.stack([t.detach().numpy() for t in lst], 0) if len(lst) else np.zeros((0,), np.float32)


def _pack_exchange(prefix, out, s, sen_w, rec_w, y, bs, br):
    out[prefix + "stop_mask"] = _stack(s[0])
    out[prefix + "stop_feat"] = _stack(s[1])
    out[prefix + "stop_prob"] = _stack(s[2])
    out[prefix + "sen_feats"] = _stack(sen_w[0])
    out[prefix + "rec_feats"] = _stack(rec_w[0])
    if sen_w[1] and sen_w[1][0] is not None:
        out[prefix + "sen_probs"] = _stack(sen_w[1])
        out[prefix + "rec_probs"] = _stack(rec_w[1])
    out[prefix + "y"] = _stack(y)
    if bs:
        out[prefix + "bs"] = _stack(bs)
        out[prefix + "br"] = _stack(br)


def make_train_case(model, name, spec):
    cfg = go.GameConfig(**spec["cfg"])
    params = go.init_params(cfg, seed=spec["seed"])
    # non-zero biases make the fixtures more discriminating than the all-zero reference init
    g = torch.Generator().manual_seed(spec["seed"] + 500)
    for a in params:
        for k, v in params[a].items():
            if k.endswith("bias") or k.endswith("bias_ih") or k.endswith("bias_hh"):
                v.add_(0.1 * torch.randn(v.shape, generator=g))
    mods = _ref_modules(model, cfg, params)
    opts = _opt(mods, cfg)
    out = {"cfg": np.array(json.dumps(cfg.as_dict())), "iters": np.array(spec["iters"])}
    for a in params:
        for k, v in params[a].items():
            out["P0/%s/%s" % (a, k)] = v.numpy().copy()
    update_src = rs.reference_update_block()
    for it in range(spec["iters"]):
        x, desc, target = go.synthetic_batch(cfg, seed=spec["seed"] * 10 + it)
        pre = "it%d/" % it
        out[pre + "x"], out[pre + "desc"], out[pre + "target"] = x.numpy(), desc.numpy(), target.numpy()
        extra = {}
        if cfg.desc_attn:
            lo, hi = spec.get("words", (1, 9))
            desc_set, lens = go.synthetic_desc_set(cfg, seed=spec["seed"] * 10 + it, min_words=lo, max_words=hi)
            out[pre + "desc_set"], out[pre + "desc_set_lens"] = desc_set.numpy(), np.array(lens, np.int32)
            extra = dict(desc_set=desc_set, desc_set_lens=lens)
        sink = []
        with rs.legacy_semantics(), rs.record_uniforms(spec["seed"] * 100 + it, sink):
            s, sen_w, rec_w, y, bs, br = model.exchange(
                mods["sender"], mods["receiver"], mods["baseline_sen"], mods["baseline_rec"],
                dict(data=x, target=target, desc=desc, train=True, break_early=not cfg.fixed_exchange, **extra))
            ns = dict(model.__dict__)
            ns.update(s=s, sen_w=sen_w, rec_w=rec_w, y=y, bs=bs, br=br, target=target,
                      sender=mods["sender"], receiver=mods["receiver"], baseline_sen=mods["baseline_sen"],
                      baseline_rec=mods["baseline_rec"], optimizer_rec=opts["receiver"],
                      optimizer_sen=opts["sender"], optimizer_bas_rec=opts["baseline_rec"],
                      optimizer_bas_sen=opts["baseline_sen"])
            exec(update_src, ns)
        steps = len(y)
        if cfg.use_binary and (cfg.flipout_sen is not None or cfg.flipout_rec is not None):
            assert cfg.flipout_sen is not None and cfg.flipout_rec is not None
            assert len(sink) == 5 * steps, (len(sink), steps)       # z, flip z, s, w, flip w
            out[pre + "u_z"] = np.stack(sink[0::5], 0)
            out[pre + "u_fz"] = np.stack(sink[1::5], 0)
            out[pre + "u_s"] = np.stack(sink[2::5], 0)
            out[pre + "u_w"] = np.stack(sink[3::5], 0)
            out[pre + "u_fw"] = np.stack(sink[4::5], 0)
        elif cfg.use_binary:
            assert len(sink) == 3 * steps, (len(sink), steps)
            out[pre + "u_z"] = np.stack(sink[0::3], 0)
            out[pre + "u_s"] = np.stack(sink[1::3], 0)
            out[pre + "u_w"] = np.stack(sink[2::3], 0)
        else:   # continuous messages: only the stop bit is sampled (model.py:420)
            assert len(sink) == steps, (len(sink), steps)
            out[pre + "u_s"] = np.stack(sink, 0)
        _pack_exchange(pre, out, s, sen_w, rec_w, y, bs, br)
        for lname in ("nll_loss", "loss_rec", "loss_sen", "loss_bas_rec", "loss_bas_sen", "loss_binary_s",
                      "loss_binary_rec", "loss_binary_sen"):
            if lname in ns and lname not in model.__dict__:
                out[pre + lname] = np.float32(float(ns[lname]))
        out[pre + "outp"] = ns["outp"].detach().numpy()
        out[pre + "logs"] = ns["logs"].detach().numpy()
        out[pre + "argmax"] = ns["argmax"].numpy().reshape(-1)
        out[pre + "ent_y_rec"] = np.array([float(e) for e in ns["ent_y_rec"]], np.float32)
        for ename in ("ent_binary_s", "ent_binary_rec", "ent_binary_sen"):
            if ename in ns and cfg.use_binary:
                out[pre + ename] = np.array([float(e) for e in ns[ename]], np.float32)
        with rs.legacy_semantics():
            pass
        # top-k accuracy line (model.py:1333-1338) restated on the reference's `dist`
        dist = ns["dist"].detach().numpy()
        top = dist.argsort()[:, -cfg.top_k_train:]
        out[pre + "accuracy"] = np.float32((top == target.numpy().reshape(-1, 1)).sum() / float(cfg.batch_size))
        # post-clip gradients of the first iteration and post-step parameters of every iteration
        keep = spec.get("store_params", go.AGENTS)
        for a, m in mods.items():
            if a not in keep:
                continue
            for k, p in m.named_parameters():
                if it == 0 and p.grad is not None:
                    out["G0/%s/%s" % (a, k)] = p.grad.detach().numpy().copy()
                out["P%d/%s/%s" % (it + 1, a, k)] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "steps", steps, "loss_rec", out[pre + "loss_rec"])


def make_eval_case(model, name, spec):
    cfg = go.GameConfig(**spec["cfg"])
    params = go.init_params(cfg, seed=spec["seed"])
    g = torch.Generator().manual_seed(spec["seed"] + 500)
    for a in params:
        for k, v in params[a].items():
            if k.endswith("bias") or k.endswith("bias_ih") or k.endswith("bias_hh"):
                v.add_(0.1 * torch.randn(v.shape, generator=g))
    params["receiver"]["s.bias"].add_(spec["s_bias"])
    params["receiver"]["s.weight"].mul_(3.0)
    mods = _ref_modules(model, cfg, params)
    rs.set_flags(model, bit_flip=spec["corrupt_region"] is not None, corrupt_region=spec["corrupt_region"])
    out = {"cfg": np.array(json.dumps(cfg.as_dict())),
           "corrupt_region": np.array(spec["corrupt_region"] or "")}
    for a in ("receiver", "sender"):
        for k, v in params[a].items():
            out["P0/%s/%s" % (a, k)] = v.numpy().copy()
    x, desc, target = go.synthetic_batch(cfg, seed=spec["seed"])
    out["x"], out["desc"], out["target"] = x.numpy(), desc.numpy(), target.numpy()
    extra = {}
    if cfg.desc_attn:
        desc_set, lens = go.synthetic_desc_set(cfg, seed=spec["seed"], min_words=1, max_words=9)
        out["desc_set"], out["desc_set_lens"] = desc_set.numpy(), np.array(lens, np.int32)
        extra = dict(desc_set=desc_set, desc_set_lens=lens)
    with rs.legacy_semantics(), torch.no_grad():
        s, sen_w, rec_w, y, bs, br = model.exchange(
            mods["sender"], mods["receiver"], None, None,
            dict(data=x, target=target, desc=desc, train=False, break_early=not cfg.fixed_exchange,
                 corrupt=spec["corrupt_region"] is not None, corrupt_region=spec["corrupt_region"], **extra))
        if cfg.fixed_exchange:
            y_masks = None
        else:
            y_masks = [torch.min(1 - m1, m2) for m1, m2 in zip(s[0][1:], s[0][:-1])]   # model.py:651-652
        outp, _ = model.get_rec_outp(y, y_masks)
    _pack_exchange("", out, s, sen_w, rec_w, y, bs, br)
    out["outp"] = outp.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "steps", len(y), "masks", [int(m.sum()) for m in s[0]])


def main():
    model = rs.load_reference()
    torch.set_num_threads(1)
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        make_train_case(model, name, spec)
    for name, spec in EVAL_CASES.items():
        if only and name not in only:
            continue
        make_eval_case(model, name, spec)


if __name__ == "__main__":
    main()
