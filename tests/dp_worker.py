"""TEST INFRASTRUCTURE — torchrun worker of the multi-GPU data-parallel parity tests (tests/test_gpu_parity.py).

Every rank drives its batch shard through the NVLink peer-memory iteration (mmg_train_step_peer) and, on a second engine,
through the NCCL all-reduce variant; rank 0 runs the single-process GLOBAL-batch CPU oracle.  Checked per iteration:
no peer wait timed out, the parameter replicas are bit-identical across ranks, peer == NCCL (within 0.25 lr with 2 ranks: the two
paths sum the batch statistics in different orders since the fused two-level statistics; within the optimizer step size otherwise), and both == oracle.

    torchrun --nproc-per-node N tests/dp_worker.py [small] [C4] [C5]

`small`: injected float64 uniforms (the reference's draw order), three configurations incl. adaptive and -desc_attn.
`C4` / `C5`: BASELINE.json configs[3] / configs[4] at their per-GPU shard shapes (64 rows, F=2048, T=10 fixed / 128 rows,
100 classes, message width 64, T=20 adaptive) with the ON-DEVICE sampler — the instantiations bench.py times; the bits every
rank drew are replayed through the oracle (u = 1 - bit reproduces the bit for any probability).
Prints "DP CHECK: PASS" on rank 0 when everything holds."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodalgame_b200 import capi, engine as eng          # noqa: E402
from oracle import game_oracle as go                         # noqa: E402
from tests import parity_util as pu                          # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = capi.load()
HEAD = dict(img_h_dim=256, baseline_hid_dim=500, sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100,
            entropy_sen=0.01, entropy_rec=0.01, top_k_train=6)
which = [a for a in sys.argv[1:]] or ["small", "C4", "C5"]
ok = True


def say(msg):
    if rank == 0:
        print(msg, flush=True)


def oracle_gap(pv, oparams):
    worst = 0.0
    for a in oparams:
        for k, v in oparams[a].items():
            if (a, k) in (("receiver", "y2.bias"), ("receiver", "d_attn.bias")):
                continue
            worst = max(worst, float((pv[a][k].detach().cpu() - v).abs().max()))
    return worst


def engines(cfg, Bl, words, params, seed=0):
    nw = int(words["desc_set"].shape[0]) if words else 0
    mk = lambda: eng.GameEngine(pu.config_from(cfg, B=Bl, batch_global=Bl * world, n_words=nw, batch_offset=rank * Bl),
                                device=dev, lib=lib, seed=seed)
    e_peer, e_nccl = mk(), mk()
    for e in (e_peer, e_nccl):
        e.load_params(params)
        if words:
            e.set_desc_set(**words)
    e_peer.enable_peer_dp()
    return e_peer, e_nccl


def compare(name, it, e_peer, e_nccl, oparams, lr, extra_ok=True):
    global ok
    err = e_peer.peer_error()
    same = torch.equal(e_peer.params, e_nccl.params)
    dmax = float((e_peer.params - e_nccl.params).abs().max())
    worst = oracle_gap(e_peer.named_views(), oparams) if rank == 0 else 0.0
    ref = e_peer.params.clone()
    dist.broadcast(ref, 0)
    rep = torch.equal(ref, e_peer.params)                    # replicas must agree bit for bit across ranks
    good = (err == 0 and rep and dmax <= (0.25 * lr if world <= 2 else 12 * lr) and worst < 3e-3 * lr * (it + 1) + 12 * lr and extra_ok)
    ok = ok and good
    say("%s it%d: peer_error=%d replicas_identical=%s peer==nccl bitwise=%s (max diff %.2e) max|param - oracle|=%.3e -> %s" % (
        name, it, err, rep, same, dmax, worst, "ok" if good else "BAD"))


if "small" in which:
    for name, kw in (("fixed", dict(fixed_exchange=True)), ("adaptive", dict(fixed_exchange=False, entropy_s=0.08)),
                     ("desc_attn", dict(fixed_exchange=True, desc_attn=True, desc_attn_dim=64))):
        Bl = 16
        B = Bl * world
        cfg = go.GameConfig(batch_size=B, img_feat_dim=512, n_classes=30, max_exchange=6, use_binary=True, **kw, **HEAD)
        params = go.init_params(cfg, seed=5)
        oparams = go.clone_params(params)
        ostate = go.new_opt_state(oparams)
        words = pu._synth_words(cfg, 5)
        e_peer, e_nccl = engines(cfg, Bl, words, params)
        lo, hi = rank * Bl, (rank + 1) * Bl
        for it in range(3):
            x, desc, target = go.synthetic_batch(cfg, seed=50 + it)
            us = go.draw_uniforms(np.random.RandomState(70 + it), cfg)
            uz, us_, uw = pu.stack_uniforms(us, cfg, B)
            sh = lambda u: u[:, lo:hi].contiguous()
            e_peer.train_step_peer(x[lo:hi], desc, target[lo:hi], uniforms=(sh(uz), sh(us_), sh(uw)))
            e_nccl.train_step_dp(x[lo:hi], desc, target[lo:hi], uniforms=(sh(uz), sh(us_), sh(uw)))
            torch.cuda.synchronize()
            if rank == 0:
                go.train_iteration(oparams, ostate, x, target, desc, cfg, us, **words)
            compare("small/" + name, it, e_peer, e_nccl, oparams, cfg.learning_rate)

SHARDS = {
    "C4": dict(batch_size=64, img_feat_dim=2048, n_classes=30, max_exchange=10, fixed_exchange=True, use_binary=True, **HEAD),
    "C5": dict(batch_size=128, img_feat_dim=2048, n_classes=100, max_exchange=20, fixed_exchange=False, use_binary=True,
               entropy_s=0.08, **dict(HEAD, sender_out_dim=64, rec_w_dim=64)),
}
for name in ("C4", "C5"):
    if name not in which:
        continue
    kw = SHARDS[name]
    Bl = kw["batch_size"]
    B = Bl * world
    cfg = go.GameConfig(**dict(kw, batch_size=B))
    T, M = cfg.max_exchange, cfg.rec_w_dim
    params = go.init_params(cfg, seed=11)
    oparams = go.clone_params(params)
    ostate = go.new_opt_state(oparams)
    e_peer, e_nccl = engines(cfg, Bl, {}, params, seed=42)
    lo, hi = rank * Bl, (rank + 1) * Bl
    for it in range(2):
        x, desc, target = go.synthetic_batch(cfg, seed=80 + it)
        e_peer.train_step_peer(x[lo:hi], desc, target[lo:hi])                 # on-device Philox draws, keyed by the global row
        torch.cuda.synchronize()
        o = e_peer.outputs()
        lv = e_peer.ws("losses", (capi.MMG_LOSS_COUNT,)).double().clone()
        dist.all_reduce(lv)                    # every rank reports its contribution to the global means
        L = dict(zip(capi.LOSS_NAMES, lv.cpu().tolist()))
        L["active_steps"] = e_peer.losses()["active_steps"]
        gathered = {}
        for key in ("sen_feats", "rec_feats", "stop_feat", "y", "sen_probs"):
            t = o[key].contiguous()
            parts = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            gathered[key] = torch.cat([p.cpu() for p in parts], 1).numpy()
        # the NCCL engine replays the very same bits (u = 1 - bit): both paths then see identical conversations
        inj = (1.0 - o["sen_feats"].double()).contiguous(), (1.0 - o["stop_feat"].double()).reshape(T, Bl).contiguous(), \
              (1.0 - o["rec_feats"].double()).contiguous()
        e_nccl.train_step_dp(x[lo:hi], desc, target[lo:hi], uniforms=inj)
        torch.cuda.synchronize()
        extra_ok = True
        if rank == 0:
            us = [(1.0 - gathered["sen_feats"][t].astype(np.float64), 1.0 - gathered["stop_feat"][t].astype(np.float64).reshape(B, 1),
                   1.0 - gathered["rec_feats"][t].astype(np.float64)) for t in range(T)]
            ex, res = go.train_iteration(oparams, ostate, x, target, desc, cfg, us)
            Tp = len(ex["y"])
            st = lambda key: np.stack([t_.detach().numpy() for t_ in ex[key]], 0)
            try:
                assert np.array_equal(gathered["sen_feats"][:Tp], st("sen_feats")) and np.array_equal(gathered["rec_feats"][:Tp], st("rec_feats"))
                pu.assert_close(name + "/y", gathered["y"][:Tp], st("y"))
                pu.assert_close(name + "/sen_probs", gathered["sen_probs"][:Tp], st("sen_probs"))
                for nm in ("nll_loss", "loss_rec", "loss_sen", "loss_bas_rec", "loss_bas_sen"):
                    pu.assert_close(name + "/" + nm, L[nm], float(res[nm].detach()))
                assert int(L["active_steps"]) == Tp, (L["active_steps"], Tp)
                # the shards of one global batch must not share their noise: rows lo..hi of different ranks differ
                if world > 1:
                    assert not np.array_equal(gathered["sen_feats"][:, :Bl], gathered["sen_feats"][:, Bl:2 * Bl]), "ranks drew identical noise"
            except AssertionError as ex_:
                print("%s it%d: %s" % (name, it, ex_), flush=True)
                extra_ok = False
        compare("%s(%d x %d rows)" % (name, world, Bl), it, e_peer, e_nccl, oparams, cfg.learning_rate, extra_ok)
        # keep the trajectories glued: continue from the oracle's parameters on every rank
        flat = {a: {k: v.clone() for k, v in d.items()} for a, d in oparams.items()}
        for a in flat:
            for k in flat[a]:
                t = flat[a][k].to(dev)
                dist.broadcast(t, 0)
                flat[a][k] = t
        e_peer.load_params(flat); e_nccl.load_params(flat)

flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
say("DP CHECK: " + ("PASS" if int(flag) == 1 else "FAIL"))
dist.destroy_process_group()
