"""CPU (emulated kernels): the reference-facing Python surface — module constructors, state_dict keys, flags,
exchange() through the autograd bridge, the mirrored loss functions and the reference's own update sequence
(model.py:1240-1330) — against the CPU oracle.  tests/test_gpu_parity.py runs the same on the CUDA library."""
import pytest

from multimodalgame_b200 import model as M
from tests import emu_util, surface_util as su


@pytest.fixture(autouse=True)
def emulated_library():
    M._LIB_OVERRIDE = emu_util.emu_library()
    M._BINDINGS.clear()
    yield
    M._LIB_OVERRIDE = None
    M._BINDINGS.clear()


@pytest.mark.parametrize("case", ["fixed_small", "adaptive_small", "continuous_t3", "fixed_t1_noent", "desc_attn_small"])
def test_reference_update_block_on_mirrored_surface(case):
    su.run_surface_case(case, "cpu")


@pytest.mark.parametrize("case", ["fixed_small", "adaptive_sgd", "desc_attn_small"])
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
