"""CPU: the oracle restatement (oracle/game_oracle.py) against golden vectors produced by the reference's own
code (tests/golden/*.npz).  This is the parity pin of the oracle."""
import numpy as np
import pytest
import torch

from oracle import game_oracle as go
from tests import golden_util as gu

TOL = dict(rtol=2e-5, atol=2e-6)


def _cmp_list(name, got_list, want, **tol):
    got = np.stack([t.detach().numpy() for t in got_list], 0)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    np.testing.assert_allclose(got, want, err_msg=name, **(tol or TOL))


@pytest.mark.parametrize("case", gu.TRAIN_CASES + gu.ATTN_TRAIN_CASES)
def test_train_iterations_match_reference(case):
    torch.set_num_threads(1)
    z, cfg = gu.load(case)
    params = gu.params_at(z, "P0")
    opt_state = go.new_opt_state(params)
    for it in range(int(z["iters"])):
        pre = "it%d/" % it
        x, desc, target = gu.batch_at(z, it)
        p_before = go.clone_params(params)
        ex, res, grads = go.train_iteration(params, opt_state, x, target, desc, cfg, gu.uniforms_at(z, it, cfg),
                                            return_grads=True, **gu.desc_set_at(z, it))
        # discrete outputs: bit-exact
        for key in ("stop_mask", "stop_feat", "sen_feats" if cfg.use_binary else None,
                    "rec_feats" if cfg.use_binary else None):
            if key is None:
                continue
            got = np.stack([t.detach().numpy() for t in ex[key]], 0)
            assert np.array_equal(got, z[pre + key]), (case, it, key)
        assert np.array_equal(res["argmax"].numpy(), z[pre + "argmax"])
        # continuous outputs
        _cmp_list("stop_prob", ex["stop_prob"], z[pre + "stop_prob"])
        # y2.bias has a mathematically-zero gradient (softmax is shift invariant) that RMSprop/Adam amplify from
        # rounding noise, so after the first update `y` is compared up to that bias.
        b_mine = float(p_before["receiver"]["y2.bias"])
        b_ref = float(gu.params_at(z, "P%d" % it)["receiver"]["y2.bias"])
        _cmp_list("y", [yy - b_mine for yy in ex["y"]], z[pre + "y"] - b_ref, rtol=2e-5, atol=1e-5)
        _cmp_list("bs", ex["bs"], z[pre + "bs"])
        _cmp_list("br", ex["br"], z[pre + "br"])
        if cfg.use_binary:
            _cmp_list("sen_probs", ex["sen_probs"], z[pre + "sen_probs"])
            _cmp_list("rec_probs", ex["rec_probs"], z[pre + "rec_probs"])
        else:
            _cmp_list("sen_feats", ex["sen_feats"], z[pre + "sen_feats"])
            _cmp_list("rec_feats", ex["rec_feats"], z[pre + "rec_feats"])
        np.testing.assert_allclose(res["outp"].detach().numpy() - b_mine, z[pre + "outp"] - b_ref, rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(res["logs"].numpy(), z[pre + "logs"], **TOL)
        for lname in ("nll_loss", "loss_rec", "loss_sen", "loss_bas_rec", "loss_bas_sen", "loss_binary_s",
                      "loss_binary_rec", "loss_binary_sen"):
            if pre + lname in z.files:
                np.testing.assert_allclose(float(res[lname]), float(z[pre + lname]), rtol=1e-4, atol=1e-5,
                                           err_msg="%s it%d %s" % (case, it, lname))
        np.testing.assert_allclose(np.array([float(e) for e in res["ent_y"]]), z[pre + "ent_y_rec"], rtol=1e-4,
                                   atol=1e-6)
        for ename, rname in (("ent_binary_s", "ent_binary_s"), ("ent_binary_rec", "ent_binary_rec"),
                             ("ent_binary_sen", "ent_binary_sen")):
            if pre + ename in z.files and rname in res:
                np.testing.assert_allclose(np.array([float(e) for e in res[rname]], np.float32), z[pre + ename],
                                           rtol=1e-4, atol=1e-6)
        assert abs(res["accuracy"] - float(z[pre + "accuracy"])) < 1e-6
        # post-step parameters
        want = gu.params_at(z, "P%d" % (it + 1))
        for a in want:
            for k, v in want[a].items():
                if (a, k) in (("receiver", "y2.bias"), ("receiver", "d_attn.bias")):   # d_attn.bias: segment softmax is shift invariant too
                    assert abs(float(params[a][k]) - float(v)) <= 12 * cfg.learning_rate * (it + 1)
                    continue
                np.testing.assert_allclose(params[a][k].numpy(), v.numpy(), rtol=1e-5, atol=2e-7,
                                           err_msg="%s it%d %s.%s" % (case, it, a, k))
        if it == 0:
            # post-clip gradients stored by the reference run
            for a in want:
                coef = min(1.0, 1.0 / (res["grad_norms"][a] + 1e-6)) if a in res["grad_norms"] else 1.0
                for k in want[a]:
                    gk = "G0/%s/%s" % (a, k)
                    if (a, k) in (("receiver", "y2.bias"), ("receiver", "d_attn.bias")):
                        continue
                    if gk in z.files:
                        assert grads[a][k] is not None, (a, k)
                        np.testing.assert_allclose((grads[a][k] * coef).numpy(), z[gk], rtol=2e-4, atol=1e-6,
                                                   err_msg=gk)
                    else:
                        assert grads.get(a, {}).get(k) is None or a not in grads, (a, k)


@pytest.mark.parametrize("case", gu.EVAL_CASES + gu.ATTN_EVAL_CASES)
def test_eval_exchange_matches_reference(case):
    torch.set_num_threads(1)
    z, cfg = gu.load(case)
    params = gu.params_at(z, "P0")
    x, desc, target = torch.from_numpy(z["x"]), torch.from_numpy(z["desc"]), torch.from_numpy(z["target"])
    region = str(z["corrupt_region"])
    mask = None
    if region:
        mask = torch.zeros(cfg.rec_w_dim)
        for r in region.split(","):
            r = r.split(":")
            idx = [int(r[0])] if len(r) == 1 else list(range(int(r[0]), int(r[1])))
            mask[idx] = 1
    with torch.no_grad():
        ex = go.exchange(params, x, desc, cfg, False, None, break_early=not cfg.fixed_exchange, corrupt_mask=mask,
                         **gu.desc_set_at(z))
    for key in ("stop_mask", "stop_feat"):
        got = np.stack([t.numpy() for t in ex[key]], 0)
        assert np.array_equal(got, z[key]), (case, key)
    _cmp_list("stop_prob", ex["stop_prob"], z["stop_prob"])
    _cmp_list("y", ex["y"], z["y"])
    if cfg.use_binary:
        for key in ("sen_feats", "rec_feats"):
            got = np.stack([t.numpy() for t in ex[key]], 0)
            assert np.array_equal(got, z[key]), (case, key)
        _cmp_list("sen_probs", ex["sen_probs"], z["sen_probs"])
        _cmp_list("rec_probs", ex["rec_probs"], z["rec_probs"])
    else:
        _cmp_list("sen_feats", ex["sen_feats"], z["sen_feats"])
        _cmp_list("rec_feats", ex["rec_feats"], z["rec_feats"])
    if cfg.fixed_exchange:
        y_masks = None
    else:
        sm = ex["stop_mask"]
        y_masks = [torch.min(1 - m1, m2) for m1, m2 in zip(sm[1:], sm[:-1])]
    outp, _ = go.get_rec_outp(ex["y"], y_masks)
    np.testing.assert_allclose(outp.numpy(), z["outp"], **TOL)
