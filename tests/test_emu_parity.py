"""CPU (no GPU): the kernel sources compiled for the thread-per-OS-thread emulator (tests/emu) against the oracle.
This checks kernel LOGIC (indexing, barrier structure, closed-form gradients) in the GPU-less build container;
the real parity gate is tests/test_gpu_parity.py, which runs the same checks through the CUDA library."""
import pytest

from tests import emu_util, parity_util as pu

CASES = ["fixed_small", "fixed_t1_noent", "continuous_t3", "adaptive_small", "adaptive_b1_adam", "adaptive_sgd",
         "flipout_small", "ignore_rec_first1", "mix_prod", "ignore_code", "desc_attn_small"]


@pytest.mark.parametrize("case", CASES)
def test_emulated_kernels_match_oracle(case):
    pu.run_train_case(case, emu_util.emu_library(), "cpu")


@pytest.mark.parametrize("case", ["eval_adaptive", "eval_adaptive_noprod", "eval_fixed_corrupt", "eval_continuous",
                                  "eval_desc_attn"])
def test_emulated_eval_matches_reference_golden(case):
    pu.run_eval_case(case, emu_util.emu_library(), "cpu")


def test_emulated_desc_attn_global_accumulators(monkeypatch):
    """-desc_attn backward with the d (d_d(word)) sums in global memory (word sets too large for shared memory)."""
    monkeypatch.setenv("MMG_ATTN_ACC_SMEM", "0")
    pu.run_train_case("desc_attn_small", emu_util.emu_library(), "cpu")


def test_emulated_synthetic_odd_dims():
    from oracle import game_oracle as go
    cfg = go.GameConfig(batch_size=7, img_feat_dim=131, img_h_dim=45, baseline_hid_dim=71, sender_out_dim=13,
                        rec_hidden=27, rec_w_dim=13, wv_dim=19, n_classes=11, max_exchange=3, fixed_exchange=False,
                        use_binary=True, entropy_s=0.05, entropy_sen=0.01, entropy_rec=0.02, top_k_train=2)
    pu.run_synth_case(cfg, emu_util.emu_library(), "cpu", iters=2, seed=6, tag="odd")


@pytest.mark.parametrize("M,fixed,binary", [(32, False, True), (64, True, True), (32, True, False)])
def test_emulated_fast_path_dims(M, fixed, binary):
    """img_h_dim=256 / rec_hidden=64 / msg_dim in {32,64} route through the specialised kernels (mmg_fast.cuh)."""
    from oracle import game_oracle as go
    cfg = go.GameConfig(batch_size=3, img_feat_dim=40, img_h_dim=256, baseline_hid_dim=24, sender_out_dim=M,
                        rec_hidden=64, rec_w_dim=M, wv_dim=12, n_classes=6, max_exchange=3, fixed_exchange=fixed,
                        use_binary=binary, entropy_s=None if fixed else 0.05, entropy_sen=0.01, entropy_rec=0.02,
                        top_k_train=2)
    pu.run_synth_case(cfg, emu_util.emu_library(), "cpu", iters=2, seed=9, tag="fast%d" % M)


@pytest.mark.parametrize("fixed", [True, False])
def test_emulated_fused_train_step_fast_dims(fixed):
    """mmg_train_step (the whole launch sequence in one C call) on the fast path."""
    from oracle import game_oracle as go
    cfg = go.GameConfig(batch_size=3, img_feat_dim=40, img_h_dim=256, baseline_hid_dim=24, sender_out_dim=32,
                        rec_hidden=64, rec_w_dim=32, wv_dim=12, n_classes=6, max_exchange=3, fixed_exchange=fixed,
                        use_binary=True, entropy_s=None if fixed else 0.05, entropy_sen=0.01, entropy_rec=0.02,
                        top_k_train=2)
    pu.run_fused_step_case(cfg, emu_util.emu_library(), "cpu", iters=2, seed=11)


def test_emulated_fast_path_flipout():
    """flipout noise (model.py:233-234,467-468) through the specialised kernels, injected flip uniforms."""
    from oracle import game_oracle as go
    cfg = go.GameConfig(batch_size=3, img_feat_dim=40, img_h_dim=256, baseline_hid_dim=24, sender_out_dim=32,
                        rec_hidden=64, rec_w_dim=32, wv_dim=12, n_classes=6, max_exchange=3, fixed_exchange=False,
                        use_binary=True, entropy_s=0.05, entropy_sen=0.01, entropy_rec=0.02, top_k_train=2,
                        flipout_sen=0.2, flipout_rec=0.1)
    pu.run_synth_case(cfg, emu_util.emu_library(), "cpu", iters=2, seed=13, tag="fastflip")


@pytest.mark.parametrize("mix,ignore", [("prod", False), ("sum", True)])
def test_emulated_fast_path_sender_variants(mix, ignore):
    """-sender_mix prod / -ignore_code (model.py:208-221) through the specialised kernels."""
    from oracle import game_oracle as go
    cfg = go.GameConfig(batch_size=3, img_feat_dim=40, img_h_dim=256, baseline_hid_dim=24, sender_out_dim=32,
                        rec_hidden=64, rec_w_dim=32, wv_dim=12, n_classes=6, max_exchange=3, fixed_exchange=True,
                        use_binary=True, entropy_sen=0.01, entropy_rec=0.02, top_k_train=2, sender_mix=mix,
                        ignore_code=ignore)
    pu.run_synth_case(cfg, emu_util.emu_library(), "cpu", iters=2, seed=15, tag="fast-" + mix)
