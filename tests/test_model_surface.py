"""CPU (emulated kernels): the reference-facing Python surface — module constructors, state_dict keys, flags,
exchange() through the autograd bridge, the mirrored loss functions and the reference's own update sequence
(model.py:1240-1330) — against the CPU oracle.  tests/test_gpu_parity.py runs the same on the CUDA library."""
import pytest

from multimodalgame_b200 import model as M
from tests import emu_util, surface_util as su


@pytest.fixture(autouse=True)
def emulated_library():
    M._LIB_OVERRIDE = emu_util.emu_library()
    M._BINDINGS.clear(); M._LAST_BINDING.clear()
    yield
    M._LIB_OVERRIDE = None
    M._BINDINGS.clear(); M._LAST_BINDING.clear()


@pytest.mark.parametrize("case", ["fixed_small", "adaptive_small", "continuous_t3", "fixed_t1_noent", "desc_attn_small", "mix_mou",
                                  "mix_mou_ignore"])
def test_reference_update_block_on_mirrored_surface(case):
    su.run_surface_case(case, "cpu")


@pytest.mark.parametrize("case", ["fixed_small", "adaptive_sgd", "adaptive_b1_adam", "desc_attn_small"])
def test_fused_train_step_behind_the_modules(case):
    su.run_train_step_case(case, "cpu")


@pytest.mark.parametrize("case", ["eval_adaptive", "eval_fixed_corrupt"])
def test_eval_dev_statistics(case):
    su.run_eval_dev_case(case, "cpu")


@pytest.mark.parametrize("case", ["fixed_small", "adaptive_b1_adam"])
def test_checkpoint_roundtrip_reference_layout(case, tmp_path):
    su.run_checkpoint_roundtrip(case, "cpu", str(tmp_path))


def test_unsupported_flags_raise():
    from tests import golden_util as gu
    z, cfg = gu.load("fixed_small")
    su.set_flags(cfg)
    with pytest.raises(NotImplementedError):
        M.Sender("layer4_2", 512, 16, 8, 8, True, True, 256, True, 1000)      # visual attention


@pytest.mark.parametrize("case,train", [("fixed_small", True), ("adaptive_small", True), ("continuous_t3", True),
                                        ("mix_prod", True), ("mix_mou", True), ("mix_mou_ignore", True), ("desc_attn_small", True), ("flipout_small", True), ("ignore_code", True),
                                        ("ignore_rec_first1", True), ("eval_adaptive", False),
                                        ("eval_desc_attn", False)])
def test_module_forwards_turn_by_turn(case, train):
    """Sender.forward / Receiver.forward / Baseline.forward (model.py:144-238, 303-477, 496-516) as single-turn calls."""
    su.run_single_turn_case(case, "cpu", train)


def test_forward_only_exchange_draws_fresh_samples():
    """model.exchange(train=True) without injected uniforms (the way INTEGRATION.md runs the reference's own update block):
    every call must consume a NEW Philox stream — the iteration counter advances in the forward itself, not in the loss."""
    import torch
    from oracle import game_oracle as go
    from tests import golden_util as gu
    z, cfg = gu.load("fixed_small")
    su.set_flags(cfg)
    params = gu.params_at(z, "P0")
    mods = su.build_modules(cfg, params, "cpu")
    x, desc, target = gu.batch_at(z, 0)
    args = dict(data=x, target=target, desc=desc, train=True)
    feats = []
    for _ in range(3):
        s, sen_w, rec_w, y, bs, br = M.exchange(mods["sender"], mods["receiver"], mods["baseline_sen"], mods["baseline_rec"], args)
        feats.append(torch.stack(sen_w[0]).clone())
    assert not torch.equal(feats[0], feats[1]) and not torch.equal(feats[1], feats[2]) and not torch.equal(feats[0], feats[2])
