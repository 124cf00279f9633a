"""CPU, world_size 2 over gloo: the data-parallel iteration — batch sharded across ranks, ONE
all-reduce of the batch statistics and ONE all-reduce of the flat gradient buffer — must reproduce the single-process
global-batch result of the oracle: `torch.std` of the REINFORCE weights, the adaptive mask counts and every mean are
over the GLOBAL batch (model.py:915,947,981).  The ranks drive the CPU-emulated build of the kernel sources
(tests/emu); on the GPU box the same host code runs over NCCL (bench.py --gpus N)."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodalgame_b200 import capi, engine as eng
from oracle import game_oracle as go
from tests import emu_util, golden_util as gu, parity_util as pu


def _worker(rank, world, case, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = emu_util.emu_library()
        z, cfg = gu.load(case)
        B = cfg.batch_size
        assert B % world == 0
        Bl = B // world
        lo, hi = rank * Bl, (rank + 1) * Bl
        e = eng.GameEngine(pu.config_from(cfg, B=Bl, batch_global=B), device="cpu", lib=lib)
        e.load_params(gu.params_at(z, "P0"))
        x, desc, target = gu.batch_at(z, 0)
        us = gu.uniforms_at(z, 0, cfg)
        uz, us_, uw = pu.stack_uniforms(us, cfg, B)
        e.train_step_dp(x[lo:hi], desc, target[lo:hi],
                        uniforms=(uz[:, lo:hi].contiguous(), us_[:, lo:hi].contiguous(), uw[:, lo:hi].contiguous()),
                        top_k=min(cfg.top_k_train, cfg.n_classes))
        losses = torch.tensor([e.losses()[n] for n in capi.LOSS_NAMES[:8]], dtype=torch.float64)
        dist.all_reduce(losses)        # loss values are rank-local contributions of global means
        out = e.outputs()
        np.savez(os.path.join(outdir, "rank%d.npz" % rank), params=e.params.numpy(), grads=e.grads.numpy(),
                 losses=losses.numpy(), sen_feats=out["sen_feats"].numpy(), rec_feats=out["rec_feats"].numpy(),
                 y=out["y"].numpy(), grad_norms=out["grad_norms"].numpy(),
                 active=np.array(e.losses()["active_steps"]))
    finally:
        dist.destroy_process_group()


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("case", ["fixed_small", "adaptive_sgd"])
def test_two_rank_iteration_matches_global_batch_oracle(case):
    emu_util.emu_library()      # build once in the parent
    world = 2
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(world, case, _free_port(), outdir), nprocs=world, join=True)
        r = [np.load(os.path.join(outdir, "rank%d.npz" % i)) for i in range(world)]
    z, cfg = gu.load(case)
    B = cfg.batch_size
    oparams = go.clone_params(gu.params_at(z, "P0"))
    x, desc, target = gu.batch_at(z, 0)
    us = gu.uniforms_at(z, 0, cfg)
    ex, res, grads = go.train_iteration(oparams, go.new_opt_state(oparams), x, target, desc, cfg, us, return_grads=True)
    Tp = len(ex["y"])
    st = lambda key: np.stack([t.detach().numpy() for t in ex[key]], 0)
    # forward: each rank holds its shard of the conversation
    for key in ("sen_feats", "rec_feats"):
        got = np.concatenate([r[i][key][:Tp] for i in range(world)], axis=1)
        assert np.array_equal(got, st(key)), key
    pu.assert_close("y", np.concatenate([r[i]["y"][:Tp] for i in range(world)], axis=1), st("y"))
    # replicas stay identical after the update
    assert np.array_equal(r[0]["params"], r[1]["params"]) and np.array_equal(r[0]["grads"], r[1]["grads"])
    assert int(r[0]["active"]) == Tp
    for i, name in enumerate(capi.LOSS_NAMES[:8]):
        if name in res:
            pu.assert_close(name, r[0]["losses"][i], float(res[name].detach()))
    # post-update parameters and clipped gradients against the global-batch oracle
    lib = emu_util.emu_library()
    e = eng.GameEngine(pu.config_from(cfg, B=B // world, batch_global=B), device="cpu", lib=lib)
    pv = e.named_views(torch.from_numpy(r[0]["params"].copy()))
    gv = e.named_views(torch.from_numpy(r[0]["grads"].copy()))
    for i, a in enumerate(capi.SEGMENTS):
        if a in res["grad_norms"]:
            pu.assert_close("grad_norm " + a, r[0]["grad_norms"][i], res["grad_norms"][a], rtol=1e-3, atol=1e-6)
    lr = cfg.learning_rate
    for a in grads:
        coef = min(1.0, 1.0 / (res["grad_norms"][a] + 1e-6))
        gmax = max([float(g.abs().max()) for g in grads[a].values() if g is not None] + [1e-12]) * coef
        for k, g in grads[a].items():
            if g is None or (a, k) == ("receiver", "y2.bias"):
                continue
            pu.assert_close("grad %s.%s" % (a, k), gv[a][k].numpy(), (g * coef).numpy(), rtol=2e-3,
                            atol=3e-5 * gmax + 1e-9)
            atol = lr * 1e-3 + 1e-7 if cfg.optim_type == "SGD" else np.where(
                (g.abs() < 1e-6).numpy(), 12 * lr, 2e-2 * lr + 1e-7)
            pu.assert_close("param %s.%s" % (a, k), pv[a][k].numpy(), oparams[a][k].numpy(), rtol=1e-5, atol=atol)
