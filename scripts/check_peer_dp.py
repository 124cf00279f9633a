"""2+ GPU check (torchrun).  With 2 ranks a + b is order-free, so the peer path must equal the NCCL path bit for bit; with more
ranks the summation orders differ and RMSprop turns rounding-level gradient differences into up to ~10 x lr of parameter
difference (same tolerance as the single-GPU parity tests).  2+ GPU check: the NVLink peer-memory data-parallel iteration against (a) the NCCL all-reduce path on the
same shards and (b) the single-process global-batch CPU oracle.  Prints PASS/FAIL on rank 0."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodalgame_b200 import capi, engine as eng
from oracle import game_oracle as go
from tests import parity_util as pu

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = capi.load()
HEAD = dict(img_h_dim=256, baseline_hid_dim=500, sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100,
            entropy_sen=0.01, entropy_rec=0.01, top_k_train=6)
ok = True
for name, kw in (("fixed", dict(fixed_exchange=True)), ("adaptive", dict(fixed_exchange=False, entropy_s=0.08)),
                 ("desc_attn", dict(fixed_exchange=True, desc_attn=True, desc_attn_dim=64))):
    Bl = 16
    B = Bl * world
    cfg = go.GameConfig(batch_size=B, img_feat_dim=512, n_classes=30, max_exchange=6, use_binary=True, **kw, **HEAD)
    params = go.init_params(cfg, seed=5)
    oparams = go.clone_params(params); ostate = go.new_opt_state(oparams)
    words = pu._synth_words(cfg, 5)
    nw = int(words["desc_set"].shape[0]) if words else 0
    e_peer = eng.GameEngine(pu.config_from(cfg, B=Bl, batch_global=B, n_words=nw), device=dev, lib=lib)
    e_nccl = eng.GameEngine(pu.config_from(cfg, B=Bl, batch_global=B, n_words=nw), device=dev, lib=lib)
    e_peer.load_params(params); e_nccl.load_params(params)
    if words:
        e_peer.set_desc_set(**words); e_nccl.set_desc_set(**words)
    e_peer.enable_peer_dp()
    lo, hi = rank * Bl, (rank + 1) * Bl
    for it in range(3):
        x, desc, target = go.synthetic_batch(cfg, seed=50 + it)
        us = go.draw_uniforms(np.random.RandomState(70 + it), cfg)
        uz, us_, uw = pu.stack_uniforms(us, cfg, B)
        sh = lambda u: u[:, lo:hi].contiguous()
        e_peer.train_step_peer(x[lo:hi], desc, target[lo:hi], uniforms=(sh(uz), sh(us_), sh(uw)))
        e_nccl.train_step_dp(x[lo:hi], desc, target[lo:hi], uniforms=(sh(uz), sh(us_), sh(uw)))
        torch.cuda.synchronize()
        go.train_iteration(oparams, ostate, x, target, desc, cfg, us, **words)
        err = e_peer.peer_error()
        same = torch.equal(e_peer.params, e_nccl.params)
        dmax = float((e_peer.params - e_nccl.params).abs().max())
        pv = e_peer.named_views()
        worst = 0.0
        for a in oparams:
            for k, v in oparams[a].items():
                if (a, k) in (("receiver", "y2.bias"), ("receiver", "d_attn.bias")):
                    continue
                worst = max(worst, float((pv[a][k].detach().cpu() - v).abs().max()))
        # replicas must agree bit for bit across ranks
        ref = e_peer.params.clone(); dist.broadcast(ref, 0)
        rep = torch.equal(ref, e_peer.params)
        good = err == 0 and rep and dmax <= (1e-7 if world <= 2 else 12 * cfg.learning_rate) and worst < 3e-3 * cfg.learning_rate * (it + 1) + 12 * cfg.learning_rate
        ok = ok and good
        if rank == 0:
            print("%s it%d: peer_error=%d replicas_identical=%s peer==nccl bitwise=%s (max diff %.2e) max|param - oracle|=%.3e -> %s" % (
                name, it, err, rep, same, dmax, worst, "ok" if good else "BAD"))
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER-DP CHECK:", "PASS" if int(flag) == 1 else "FAIL")
dist.destroy_process_group()
