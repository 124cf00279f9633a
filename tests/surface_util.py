"""Drives the reference-facing Python surface (multimodalgame_b200/model.py: Sender / Receiver / Baseline modules,
exchange(), the loss functions, train_step()) exactly the way the reference's run() does (model.py:1240-1330) and
compares with the CPU oracle.  Used by the CPU test (emulated kernels) and the GPU test (CUDA library)."""
import numpy as np
import torch
import torch.nn.functional as F

from multimodalgame_b200 import model as M
from oracle import game_oracle as go
from tests import golden_util as gu, parity_util as pu

# state_dict keys of the reference's modules (verified against the reference classes)
REF_KEYS = {
    "sender": ["code_bias", "image_layer.weight", "image_layer.bias", "code_layer.weight", "code_layer.bias",
               "binary_layer.weight", "binary_layer.bias"],
    "receiver": ["rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh", "w_h.weight", "w_h.bias", "w_d.weight",
                 "w.weight", "w.bias", "y1.weight", "y1.bias", "y2.weight", "y2.bias", "s.weight", "s.bias"],
    "baseline": ["linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias"],
}
ATTN_KEYS = ["d_d.weight", "d_d.bias", "d_h.weight", "d_h.bias", "d_attn.weight", "d_attn.bias"]   # -desc_attn, model.py:267-271
ZERO_GRAD = (("receiver", "y2.bias"), ("receiver", "d_attn.bias"))    # softmax shift invariance: rounding noise only

_flags_defined = False


def set_flags(cfg):
    """Parse a reference-style command line (single-dash gflags syntax, model.py:1639-1741)."""
    global _flags_defined
    if not _flags_defined:
        M.flags()
        _flags_defined = True
    argv = ["model.py", "-max_exchange", str(cfg.max_exchange), "-learning_rate", repr(cfg.learning_rate),
            "-optim_type", cfg.optim_type, "-top_k_train", str(cfg.top_k_train), "-first_rec", repr(cfg.first_rec),
            "-img_feat_dim", str(cfg.img_feat_dim), "-img_h_dim", str(cfg.img_h_dim), "-baseline_hid_dim",
            str(cfg.baseline_hid_dim), "-sender_out_dim", str(cfg.sender_out_dim), "-rec_w_dim", str(cfg.rec_w_dim),
            "-rec_hidden", str(cfg.rec_hidden), "-wv_dim", str(cfg.wv_dim), "-batch_size", str(cfg.batch_size)]
    argv.append("-fixed_exchange" if cfg.fixed_exchange else "-nofixed_exchange")
    argv.append("-use_binary" if cfg.use_binary else "-nouse_binary")
    argv.append("-s_prob_prod" if cfg.s_prob_prod else "-nos_prob_prod")
    if getattr(cfg, "desc_attn", False):
        argv += ["-desc_attn", "-desc_attn_dim", str(cfg.desc_attn_dim)]
    else:
        argv.append("-nodesc_attn")
    for name in ("entropy_s", "entropy_sen", "entropy_rec", "flipout_sen", "flipout_rec"):
        if getattr(cfg, name, None) is not None:
            argv += ["-" + name, repr(getattr(cfg, name))]
    argv += ["-sender_mix", getattr(cfg, "sender_mix", "sum")]
    argv.append("-ignore_code" if getattr(cfg, "ignore_code", False) else "-noignore_code")
    argv.append("-ignore_receiver" if getattr(cfg, "ignore_receiver", False) else "-noignore_receiver")
    M.FLAGS.unparse_flags()
    for name in ("entropy_s", "entropy_sen", "entropy_rec", "flipout_sen", "flipout_rec"):
        M.FLAGS[name].value = None
    M.FLAGS(argv)
    M.default_flags(argv)


def build_modules(cfg, params, device):
    """The four modules with the reference's constructor signatures (model.py:1014-1064)."""
    sender = M.Sender("avgpool_512", cfg.img_feat_dim, cfg.img_h_dim, cfg.rec_w_dim, cfg.sender_out_dim, cfg.use_binary,
                      False, 0, False, 0)
    receiver = M.Receiver(cfg.sender_out_dim, cfg.wv_dim, cfg.rec_hidden, 1, cfg.rec_w_dim, 1, cfg.use_binary)
    baseline_sen = M.Baseline(cfg.baseline_hid_dim, cfg.img_h_dim, cfg.rec_w_dim, 0)
    baseline_rec = M.Baseline(cfg.baseline_hid_dim, 0, cfg.sender_out_dim, cfg.rec_hidden)
    mods = dict(sender=sender, receiver=receiver, baseline_sen=baseline_sen, baseline_rec=baseline_rec)
    want = list(REF_KEYS["sender"])
    if getattr(cfg, "sender_mix", "sum") == "mou" and getattr(cfg, "ignore_code", False):
        want.insert(1, "code_bias_mou")                  # direct parameters come first in state_dict order (model.py:73-74)
    assert list(sender.state_dict().keys()) == want
    assert list(receiver.state_dict().keys()) == REF_KEYS["receiver"] + (ATTN_KEYS if getattr(cfg, "desc_attn", False) else [])
    assert list(baseline_sen.state_dict().keys()) == REF_KEYS["baseline"]
    for a, m in mods.items():
        if a in params:
            m.load_state_dict(params[a])          # reference checkpoints load by key name
        m.to(device)
    return mods


def reference_update_block(mods, cfg, x, desc, target, uniforms, words=None):
    """model.py:1240-1330 written against the mirrored names; returns the loss dict (gradients land in .grad)."""
    fl = M.FLAGS
    exchange_args = dict(data=x, target=target, desc=desc, train=True, break_early=not fl.fixed_exchange,
                         uniforms=uniforms, **(words or {}))
    s, sen_w, rec_w, y, bs, br = M.exchange(mods["sender"], mods["receiver"], mods["baseline_sen"], mods["baseline_rec"],
                                            exchange_args)
    s_masks, s_feats, s_probs = s
    sen_feats, sen_probs = sen_w
    rec_feats, rec_probs = rec_w
    if fl.fixed_exchange:                                                     # model.py:1248-1262
        binary_s_masks = binary_rec_masks = binary_sen_masks = bas_rec_masks = bas_sen_masks = y_masks = None
    else:
        binary_s_masks = s_masks[:-1]
        binary_rec_masks = s_masks[1:-1]
        binary_sen_masks = s_masks[:-1]
        bas_rec_masks = s_masks[:-1]
        bas_sen_masks = s_masks[:-1]
        y_masks = [torch.min(1 - m2, m1) for m1, m2 in zip(s_masks[:-1], s_masks[1:])]
    outp, _ = M.get_rec_outp(y, y_masks)                                      # model.py:1264
    dist = F.log_softmax(outp, dim=1)
    nll_loss = F.nll_loss(dist, target)
    logs = M.loglikelihood(dist.detach(), target.view(-1, 1))
    out = dict(nll_loss=nll_loss, steps=len(y), y=y, sen_feats=sen_feats, rec_feats=rec_feats, sen_probs=sen_probs,
               s_masks=s_masks)
    for m in mods.values():
        m.zero_grad()
    if cfg.use_binary:
        loss_binary_s = None
        if not fl.fixed_exchange:
            loss_binary_s, _ = M.multistep_loss_binary(s_feats, s_probs, logs, br, binary_s_masks, fl.entropy_s)
        if len(rec_feats[:-1]) > 0:                                            # model.py:1284-1289
            loss_binary_rec, _ = M.multistep_loss_binary(rec_feats[:-1], rec_probs[:-1], logs, br[:-1], binary_rec_masks,
                                                         fl.entropy_rec)
        else:
            loss_binary_rec = torch.zeros((), device=x.device)
        loss_binary_sen, _ = M.multistep_loss_binary(sen_feats, sen_probs, logs, bs, binary_sen_masks, fl.entropy_sen)
        loss_bas_rec = M.multistep_loss_bas(br, logs, bas_rec_masks)
        loss_bas_sen = M.multistep_loss_bas(bs, logs, bas_sen_masks)
        loss_rec = nll_loss + loss_binary_rec + (loss_binary_s if loss_binary_s is not None else 0)
        loss_sen = loss_binary_sen
        loss_rec.backward()                                                    # model.py:1309, 1316, 1322, 1328
        loss_sen.backward()
        loss_bas_rec.backward()
        loss_bas_sen.backward()
        out.update(loss_rec=loss_rec, loss_sen=loss_sen, loss_bas_rec=loss_bas_rec, loss_bas_sen=loss_bas_sen)
    else:
        nll_loss.backward()
        out.update(loss_rec=nll_loss)
    return out


def run_surface_case(case, device, iters=1):
    z, cfg = gu.load(case)
    set_flags(cfg)
    params = gu.params_at(z, "P0")
    full = go.init_params(cfg, seed=1)            # fixtures may store a subset of the agents
    for a in full:
        if a not in params:
            params[a] = full[a]
    oparams = go.clone_params(params)
    mods = build_modules(cfg, params, device)
    B = cfg.batch_size
    x, desc, target = gu.batch_at(z, 0)
    us = gu.uniforms_at(z, 0, cfg)
    words = gu.desc_set_at(z, 0)
    ex, res, grads = go.train_iteration(oparams, go.new_opt_state(oparams), x, target, desc, cfg, us, return_grads=True,
                                        **words)
    uni = tuple(u.to(device) for u in pu.stack_uniforms(us, cfg, B))
    out = reference_update_block(mods, cfg, x.to(device), desc.to(device), target.to(device), uni, words)
    Tp = len(ex["y"])
    assert out["steps"] == Tp, (out["steps"], Tp)
    st = lambda key: np.stack([t.detach().numpy() for t in ex[key]], 0)
    got = lambda lst: np.stack([t.detach().cpu().numpy() for t in lst], 0)
    pu.assert_close(case + "/y", got(out["y"]), st("y"))
    if cfg.use_binary:
        assert np.array_equal(got(out["sen_feats"]), st("sen_feats"))
        assert np.array_equal(got(out["rec_feats"]), st("rec_feats"))
        pu.assert_close(case + "/sen_probs", got(out["sen_probs"]), st("sen_probs"))
    assert len(out["s_masks"]) == Tp + 1 and float(out["s_masks"][-1].sum()) == 0.0      # model.py:870
    for name in ("nll_loss", "loss_rec", "loss_sen", "loss_bas_rec", "loss_bas_sen"):
        if name in out and name in res:
            pu.assert_close(case + "/" + name, float(out[name]), float(res[name]))
    for a, mod in mods.items():
        if a not in grads:
            continue
        gmax = max([float(g.abs().max()) for g in grads[a].values() if g is not None] + [1e-12])
        for k, p in mod.named_parameters():
            g = grads[a].get(k)
            if g is None or (a, k) in ZERO_GRAD:
                continue
            assert p.grad is not None, (a, k)
            pu.assert_close("%s/grad %s.%s" % (case, a, k), p.grad.detach().cpu().numpy(), g.numpy(), rtol=2e-3,
                            atol=2e-5 * gmax + 1e-9)
    return out


def run_train_step_case(case, device):
    """model.train_step(): the fused iteration behind the reference's module objects; module parameters (views of the
    engine's flat buffer) must equal the oracle's post-step parameters, and .grad must hold the clipped gradients."""
    z, cfg = gu.load(case)
    set_flags(cfg)
    params = gu.params_at(z, "P0")
    full = go.init_params(cfg, seed=1)
    for a in full:
        if a not in params:
            params[a] = full[a]
    oparams = go.clone_params(params)
    mods = build_modules(cfg, params, device)
    x, desc, target = gu.batch_at(z, 0)
    us = gu.uniforms_at(z, 0, cfg)
    words = gu.desc_set_at(z, 0)
    ex, res, grads = go.train_iteration(oparams, go.new_opt_state(oparams), x, target, desc, cfg, us, return_grads=True,
                                        **words)
    uni = tuple(u.to(device) for u in pu.stack_uniforms(us, cfg, cfg.batch_size))
    eng = M.train_step(mods["sender"], mods["receiver"], mods["baseline_sen"], mods["baseline_rec"],
                       dict(data=x.to(device), target=target.to(device), desc=desc.to(device), train=True, uniforms=uni,
                            **words))
    L = eng.losses()
    for name in ("nll_loss", "loss_rec", "loss_sen", "loss_bas_rec", "loss_bas_sen"):
        if name in res:
            pu.assert_close(case + "/" + name, L[name], float(res[name].detach()))
    lr = cfg.learning_rate
    for a, mod in mods.items():
        if a not in oparams:
            continue
        for k, p in mod.named_parameters():
            if (a, k) in ZERO_GRAD:
                continue
            g = grads.get(a, {}).get(k)
            atol = 2e-2 * lr + 1e-7
            if cfg.optim_type == "SGD":
                atol = lr * 1e-3 + 1e-7
            elif g is not None:
                atol = np.where((g.abs() < 1e-5).numpy(), 12 * lr, atol)     # same noise-level rule as parity_util
            pu.assert_close("%s/param %s.%s" % (case, a, k), p.detach().cpu().numpy(), oparams[a][k].numpy(), rtol=1e-5,
                            atol=atol)
    return eng


def run_eval_dev_case(case, device, top_k=2):
    """model.eval_dev() on a reference-generated eval fixture (two identical batches) against the statistics recomputed in
    NumPy from the fixture's arrays exactly as model.py:648-718 does."""
    z, cfg = gu.load(case)
    set_flags(cfg)
    region = str(z["corrupt_region"])
    M.FLAGS.bit_flip = bool(region)
    M.FLAGS.corrupt_region = region or None
    M.FLAGS.conf_mat = ""
    params = gu.params_at(z, "P0")
    full = go.init_params(cfg, seed=1)
    for a in full:
        if a not in params:
            params[a] = full[a]
    mods = build_modules(cfg, params, device)
    x, desc, target = torch.from_numpy(z["x"]), torch.from_numpy(z["desc"]), torch.from_numpy(z["target"])
    batches = [{"target": target, M.FLAGS.img_feat: x}, {"target": target, M.FLAGS.img_feat: x}]
    acc, extra = M.eval_dev(batches, cfg.batch_size, 0, False, device != "cpu", top_k, mods["sender"], mods["receiver"],
                            dict(desc=desc), None, None)
    # expected, from the reference's own outputs
    dist = torch.log_softmax(torch.from_numpy(z["outp"]), dim=1).numpy()
    topk = np.argsort(dist, axis=1)[:, -top_k:]
    tgt = z["target"].reshape(-1)
    correct = float((topk == tgt[:, None]).sum())
    pu.assert_close(case + "/accuracy", acc, correct / cfg.batch_size)
    lengths = z["stop_feat"].reshape(z["stop_feat"].shape[0], -1).sum(0)
    pu.assert_close(case + "/conv_mean", extra["conversation_lengths_mean"], lengths.mean())
    pu.assert_close(case + "/conv_std", extra["conversation_lengths_std"], lengths.std())
    for key, name in (("sen_feats", "hamming_sen_mean"), ("rec_feats", "hamming_rec_mean")):
        msgs = z[key]
        prev = np.concatenate([np.zeros_like(msgs[:1]), msgs[:-1]], 0)
        pu.assert_close(case + "/" + name, extra[name], np.abs(msgs - prev).sum(2).mean(1).sum() / msgs.shape[0], rtol=1e-4,
                        atol=1e-4)
    pred = dist.argmax(1)
    cm = np.zeros((cfg.n_classes, cfg.n_classes), dtype=np.int64)
    for t, p in zip(tgt, pred):
        cm[t, p] += 2
    assert np.array_equal(extra["confusion_matrix"], cm)
    M.FLAGS.bit_flip = False
    M.FLAGS.corrupt_region = None


def run_checkpoint_roundtrip(case, device, tmpdir):
    """train 2 iterations -> torch_save (reference dictionary layout) -> fresh modules -> torch_load -> third iteration must
    equal the uninterrupted run bit for bit (module parameters AND fused optimizer state travel through the file)."""
    import os
    z, cfg = gu.load(case)
    set_flags(cfg)
    params = gu.params_at(z, "P0")

    def fresh():
        M._BINDINGS.clear(); M._LAST_BINDING.clear()
        mods = build_modules(cfg, params, device)
        return mods

    def step(mods, it):
        x, desc, target = gu.batch_at(z, it % int(z["iters"]))
        us = gu.uniforms_at(z, it % int(z["iters"]), cfg)
        uni = tuple(u.to(device) for u in pu.stack_uniforms(us, cfg, cfg.batch_size))
        return M.train_step(mods["sender"], mods["receiver"], mods["baseline_sen"], mods["baseline_rec"],
                            dict(data=x.to(device), target=target.to(device), desc=desc.to(device), train=True, uniforms=uni))

    names = dict(receiver="optimizer_rec", sender="optimizer_sen", baseline_rec="optimizer_bas_rec", baseline_sen="optimizer_bas_sen")
    a = fresh()
    for it in range(3):
        ea = step(a, it)
    want = {k: {n: p.detach().cpu().clone() for n, p in m.named_parameters()} for k, m in a.items()}
    b = fresh()
    for it in range(2):
        eb = step(b, it)
    path = os.path.join(tmpdir, "ckpt.pt")
    M.torch_save(path, dict(step=2, best_dev_acc=0.5), b, {names[k]: M.FusedOptimizer(eb, k) for k in names}, -1)
    ck = torch.load(path, weights_only=False)
    assert sorted(ck.keys()) == ["data", "models", "optimizers"]                        # misc.py:68-72
    assert list(ck["models"]["receiver"].keys()) == REF_KEYS["receiver"]
    c = fresh()
    ec = step(c, 0)                    # binds modules to a new engine (state gets overwritten by the load)
    data = M.torch_load(path, c, {names[k]: M.FusedOptimizer(ec, k) for k in names})
    assert data["step"] == 2
    ec = step(c, 2)
    for k, m in c.items():
        for n, p in m.named_parameters():
            assert torch.equal(p.detach().cpu(), want[k][n]), (k, n)


def run_single_turn_case(case, device, train=True):
    """Sender.forward / Receiver.forward / Baseline.forward driven turn by turn exactly as the reference's exchange() loop
    does (model.py:801-867), against the oracle's conversation on the same inputs and injected uniforms."""
    z, cfg = gu.load(case)
    set_flags(cfg)
    params = gu.params_at(z, "P0")
    full = go.init_params(cfg, seed=1)
    for a in full:
        if a not in params:
            params[a] = full[a]
    mods = build_modules(cfg, params, device)
    sender, receiver = mods["sender"], mods["receiver"]
    if "iters" in z.files:
        x, desc, target = gu.batch_at(z, 0)
        us = gu.uniforms_at(z, 0, cfg) if train else None
        words = gu.desc_set_at(z, 0)
    else:                                     # eval fixtures store one batch
        x, desc, target = torch.from_numpy(z["x"]), torch.from_numpy(z["desc"]), torch.from_numpy(z["target"])
        us, words = None, {}
        if "desc_set" in z.files:
            words = dict(desc_set=torch.from_numpy(z["desc_set"]), desc_set_lens=[int(v) for v in z["desc_set_lens"]])
    if train and us is None:
        us = go.draw_uniforms(np.random.RandomState(5), cfg)
    ex = go.exchange(go.clone_params(params), x, desc, cfg, train, uniforms=us, break_early=not cfg.fixed_exchange, **words)
    for m in mods.values():
        m.train() if train else m.eval()
    sender.reset_state(); receiver.reset_state()
    B = x.shape[0]
    xd, dd = x.to(device), desc.to(device)
    w = torch.full((B, cfg.rec_w_dim), float(cfg.first_rec), device=device)            # model.py:786
    wkw = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in words.items()}
    for t in range(len(ex["y"])):
        u = [None if a is None else torch.from_numpy(np.asarray(a, dtype=np.float64)) for a in us[t]] if train else [None] * 5
        u = u + [None] * (5 - len(u))
        z_r = w
        zb, zp = sender.forward(xd, z_r, None, t, uniforms=(u[0], u[3]) if train else None)
        (s, sp), (w, wp), y = receiver.forward(zb, dd, uniforms=(u[1], u[2], u[4]) if train else None, **wkw)
        tag = "%s/t%d/" % (case, t)
        npy = lambda v: v.detach().cpu().numpy()
        pu.assert_close(tag + "y", npy(y), ex["y"][t].detach().numpy())
        pu.assert_close(tag + "s_prob", npy(sp), ex["stop_prob"][t].detach().numpy())
        pu.assert_close(tag + "h_x", npy(sender.h_x), ex["h_x"].detach().numpy())
        pu.assert_close(tag + "h_z", npy(receiver.h_z), ex["h_z"][t].detach().numpy())
        pu.assert_close(tag + "h_w", npy(receiver.h_w), ex["h_w"][t].detach().numpy())
        if cfg.use_binary:
            pu.assert_close(tag + "sen_probs", npy(zp), ex["sen_probs"][t].detach().numpy())
            pu.assert_close(tag + "rec_probs", npy(wp), ex["rec_probs"][t].detach().numpy())
            assert np.array_equal(npy(zb), ex["sen_feats"][t].numpy()), tag + "sender bits"
            assert np.array_equal(npy(w), ex["rec_feats"][t].numpy()), tag + "receiver bits"
            assert np.array_equal(npy(s), ex["stop_feat"][t].numpy()), tag + "stop bits"
        else:
            assert zp is None and wp is None
            pu.assert_close(tag + "sen_feats", npy(zb), ex["sen_feats"][t].detach().numpy())
            pu.assert_close(tag + "rec_feats", npy(w), ex["rec_feats"][t].detach().numpy())
        if train:
            bs = mods["baseline_sen"].forward(sender.h_x, z_r, None)                    # model.py:835-836
            br = mods["baseline_rec"].forward(None, zb, receiver.h_z)                   # model.py:842-843
            pu.assert_close(tag + "bs", npy(bs), ex["bs"][t].detach().numpy())
            pu.assert_close(tag + "br", npy(br), ex["br"][t].detach().numpy())
    return len(ex["y"])
