"""ctypes binding of the C-ABI declared in include/mmg_b200.h (libmmg_b200.so).

The product path loads exactly one library: the CUDA build next to this file.  There is no CPU fallback: if the
library is missing, or no CUDA device is visible, `load()` raises.  (tests/emu builds the same kernel sources for a
CPU emulator and hands its own handle to `Library(...)` explicitly; nothing here ever looks for it.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libmmg_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

MMG_ABI_VERSION = 6          # include/mmg_b200.h; a library built from other headers is refused at load time
MMG_P_COUNT = 37
MMG_SEG_COUNT = 4
MMG_LOSS_COUNT = 16

# (agent, state_dict key) for every MMG_P_* index, in enum order (include/mmg_b200.h)
PARAM_NAMES = [
    ("receiver", "rnn.weight_ih"), ("receiver", "rnn.weight_hh"), ("receiver", "rnn.bias_ih"),
    ("receiver", "rnn.bias_hh"), ("receiver", "w_h.weight"), ("receiver", "w_h.bias"), ("receiver", "w_d.weight"),
    ("receiver", "w.weight"), ("receiver", "w.bias"), ("receiver", "y1.weight"), ("receiver", "y1.bias"),
    ("receiver", "y2.weight"), ("receiver", "y2.bias"), ("receiver", "s.weight"), ("receiver", "s.bias"),
    ("receiver", "d_d.weight"), ("receiver", "d_d.bias"), ("receiver", "d_h.weight"), ("receiver", "d_h.bias"),
    ("receiver", "d_attn.weight"), ("receiver", "d_attn.bias"),      # -desc_attn only, zero-sized otherwise
    ("sender", "code_bias"), ("sender", "code_bias_mou"),            # code_bias_mou: -sender_mix mou -ignore_code only, zero-sized otherwise
    ("sender", "image_layer.weight"), ("sender", "image_layer.bias"),
    ("sender", "code_layer.weight"), ("sender", "code_layer.bias"), ("sender", "binary_layer.weight"),
    ("sender", "binary_layer.bias"),
    ("baseline_rec", "linear1.weight"), ("baseline_rec", "linear1.bias"), ("baseline_rec", "linear2.weight"),
    ("baseline_rec", "linear2.bias"),
    ("baseline_sen", "linear1.weight"), ("baseline_sen", "linear1.bias"), ("baseline_sen", "linear2.weight"),
    ("baseline_sen", "linear2.bias"),
]
SEGMENTS = ("receiver", "sender", "baseline_rec", "baseline_sen")
LOSS_NAMES = ("nll_loss", "loss_rec", "loss_sen", "loss_bas_rec", "loss_bas_sen", "loss_binary_s", "loss_binary_rec",
              "loss_binary_sen", "topk_correct", "active_steps")
OPTIM = {"RMSprop": 0, "Adam": 1, "SGD": 2}
SENDER_MIX = {"sum": 0, "prod": 1, "mou": 2}


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "batch_global", "img_feat_dim", "img_h_dim", "msg_dim", "rec_hidden", "n_classes", "wv_dim",
        "baseline_hid", "max_exchange", "use_binary", "fixed_exchange", "s_prob_prod", "optim_type",
        "has_entropy_s", "has_entropy_sen", "has_entropy_rec")] + [
        (n, C.c_float) for n in ("entropy_s", "entropy_sen", "entropy_rec", "first_rec", "learning_rate", "max_norm")
    ] + [("ignore_receiver", C.c_int32), ("has_flipout_sen", C.c_int32), ("has_flipout_rec", C.c_int32),
         ("flipout_dev", C.c_int32), ("flipout_sen", C.c_float), ("flipout_rec", C.c_float), ("sender_mix", C.c_int32),
         ("ignore_code", C.c_int32), ("desc_attn", C.c_int32), ("desc_attn_dim", C.c_int32), ("n_words", C.c_int32),
         ("batch_offset", C.c_int32)]


class ParamLayout(C.Structure):
    _fields_ = [("offset", C.c_int64 * MMG_P_COUNT), ("rows", C.c_int32 * MMG_P_COUNT),
                ("cols", C.c_int32 * MMG_P_COUNT), ("segment", C.c_int32 * MMG_P_COUNT),
                ("seg_begin", C.c_int64 * (MMG_SEG_COUNT + 1)), ("total", C.c_int64)]


_WS_FIELDS = ("total_bytes", "sen_feats", "sen_probs", "rec_feats", "rec_probs", "stop_feat", "stop_prob", "y",
              "stop_mask", "bs", "br", "h_x", "h_z", "h_w", "losses", "ystep", "outp", "logs", "argmax", "stats",
              "stats_count", "grad_norms", "g_sen_probs", "g_rec_probs", "g_stop_prob", "g_outp", "g_bs", "g_br",
              "rng_state", "opt_counters")


class WorkspaceLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in _WS_FIELDS]


class Inputs(C.Structure):
    _fields_ = [("d_x", C.c_void_p), ("d_desc", C.c_void_p), ("d_target", C.c_void_p), ("d_u_sen", C.c_void_p),
                ("d_u_stop", C.c_void_p), ("d_u_rec", C.c_void_p), ("d_corrupt_mask", C.c_void_p),
                ("d_h0", C.c_void_p), ("top_k", C.c_int32), ("train", C.c_int32), ("d_u_flip_sen", C.c_void_p),
                ("d_u_flip_rec", C.c_void_p), ("d_desc_set", C.c_void_p), ("d_desc_set_lens", C.c_void_p),
                ("h_losses_out", C.c_void_p)]


MMG_MAX_PEERS = 8


class Peers(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("d_send", C.c_void_p * MMG_MAX_PEERS),
                ("d_recv", C.c_void_p * MMG_MAX_PEERS), ("d_stats", C.c_void_p * MMG_MAX_PEERS),
                ("d_norms", C.c_void_p * MMG_MAX_PEERS), ("d_flags", C.c_void_p * MMG_MAX_PEERS), ("d_error", C.c_void_p),
                ("d_send_mc", C.c_void_p)]


class MmgError(RuntimeError):
    pass


class Library(object):
    """Typed view of a loaded libmmg shared object."""

    SYMBOLS = ("mmg_abi_version", "mmg_last_error", "mmg_device_count", "mmg_param_layout_get",
               "mmg_workspace_layout_get", "mmg_workspace_init", "mmg_exchange_forward", "mmg_loss", "mmg_backward",
               "mmg_grad_norm", "mmg_clip_update", "mmg_train_step", "mmg_train_step_host", "mmg_host_prefetch", "mmg_train_step_staged", "mmg_peer_buffer_layout", "mmg_train_step_peer", "mmg_launch_count",
               "mmg_launch_count_reset", "mmg_sender_forward", "mmg_receiver_forward", "mmg_baseline_forward",
               "mmg_debug_kernel_times", "mmg_debug_trace")

    def __init__(self, path):
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        vp, i64, f32 = C.c_void_p, C.c_int64, C.c_float
        cfgp = C.POINTER(Config)
        inp = C.POINTER(Inputs)
        d.mmg_abi_version.restype = C.c_int
        d.mmg_last_error.restype = C.c_char_p
        d.mmg_device_count.restype = C.c_int
        d.mmg_launch_count.restype = C.c_int
        d.mmg_launch_count_reset.restype = None
        d.mmg_param_layout_get.argtypes = [cfgp, C.POINTER(ParamLayout)]
        d.mmg_workspace_layout_get.argtypes = [cfgp, C.POINTER(WorkspaceLayout)]
        d.mmg_workspace_init.argtypes = [cfgp, vp, C.c_uint64, vp]
        d.mmg_exchange_forward.argtypes = [cfgp, vp, inp, vp, vp]
        d.mmg_loss.argtypes = [cfgp, vp, inp, vp, C.c_int, vp]
        d.mmg_backward.argtypes = [cfgp, vp, inp, vp, vp, vp]
        d.mmg_grad_norm.argtypes = [cfgp, vp, vp, vp]
        d.mmg_clip_update.argtypes = [cfgp, vp, vp, vp, vp, i64, f32, vp, vp]
        d.mmg_train_step.argtypes = [cfgp, vp, vp, vp, vp, i64, inp, vp, vp]
        d.mmg_train_step_host.argtypes = [cfgp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, inp, vp, vp, vp]
        d.mmg_host_prefetch.argtypes = [cfgp, vp, vp, vp, vp, vp, vp, vp]
        d.mmg_train_step_staged.argtypes = [cfgp, vp, vp, vp, vp, i64, inp, vp, vp, vp, vp, vp]
        i64p = C.POINTER(C.c_int64)
        d.mmg_peer_buffer_layout.argtypes = [cfgp, i64p, i64p, i64p, i64p, i64p, i64p]
        d.mmg_train_step_peer.argtypes = [cfgp, vp, vp, vp, vp, i64, inp, vp, C.POINTER(Peers), vp]
        i32, u64 = C.c_int32, C.c_uint64
        d.mmg_sender_forward.argtypes = [cfgp, vp, i32, vp, vp, i32, i32, vp, vp, u64, u64, vp, vp, vp, vp]
        d.mmg_receiver_forward.argtypes = [cfgp, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, u64, u64,
                                           vp, vp, vp, vp, vp, vp, vp]
        d.mmg_baseline_forward.argtypes = [cfgp, vp, i32, i32, vp, i32, vp, i32, vp, i32, vp, vp]
        d.mmg_debug_kernel_times.argtypes = [C.c_char_p, i32]
        d.mmg_debug_trace.argtypes = [vp, i32]
        for name in self.SYMBOLS:
            fn = getattr(d, name)
            if name not in ("mmg_last_error", "mmg_launch_count_reset"):
                fn.restype = C.c_int
        got = int(d.mmg_abi_version())
        if got != MMG_ABI_VERSION:
            raise MmgError("%s implements ABI version %d, this binding expects %d (mmg_config / entry points differ): "
                           "rebuild it with __graft_entry__.build(force=True)" % (path, got, MMG_ABI_VERSION))

    def check(self, rc, what):
        if rc < 0:
            raise MmgError("%s failed (%d): %s" % (what, rc, self.dll.mmg_last_error().decode()))
        return rc

    def call(self, name, *args):
        return self.check(getattr(self.dll, name)(*args), name)


_LIB = None


def load():
    """Load the CUDA library.  Raises if it has not been built or no CUDA device is visible."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise MmgError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        lib = Library(LIB_PATH)
        n = lib.dll.mmg_device_count()
        if n <= 0:
            raise MmgError("no CUDA device visible (mmg_device_count=%d): %s" % (n, lib.dll.mmg_last_error().decode()))
        _LIB = lib
    return _LIB
