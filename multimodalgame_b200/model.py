"""Host-side mirror of the reference's operator surface for the hot path (`/root/reference/model.py`).

Same names, constructor/`forward` signatures, flag names and `state_dict` keys as the reference, so that the
reference's `run()` loop (model.py:1190-1339) can call into this module unchanged:

    Sender / Receiver / Baseline      model.py:49-238 / 241-477 / 480-516
    exchange(...)                     model.py:725-876
    eval_dev(...)                     model.py:580-722 (batches as dicts; statistics accumulated on the device)
    get_rec_outp, calculate_loss_binary, multistep_loss_binary, calculate_loss_bas, multistep_loss_bas,
    build_inp, flipout, loglikelihood model.py:519-577, 879-988
    flags(), default_flags(), Fixed/Adaptive presets       model.py:1605-1810

Everything the conversation computes runs in libmmg_b200.so (hand-written sm_100a kernels behind the C-ABI in
include/mmg_b200.h).  Two ways in:

  * `exchange()` — drop-in.  Returns the reference's tuple of lists; in train mode the probabilities / scores /
    baseline values carry autograd history through one custom Function node per module whose backward is the fused
    backward kernel sequence, so `loss.backward()` + `torch.optim` work exactly as in the reference's loop.
  * `train_step()` — the whole iteration (conversation, the five losses, backward, 4x clip + optimizer step) fused on
    the device with no host round trip.  This is what `bench.py` measures.

There is no CPU path: without the CUDA library and a GPU these functions raise.
"""
import ctypes as C
import math
import os
import sys
import weakref

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import capi
from . import engine as _engine

try:
    from absl import flags as gflags
except ImportError:  # pragma: no cover
    import gflags

FLAGS = gflags.FLAGS

_LIB_OVERRIDE = None   # tests hand the emulated build in here; the product never sets it
_SAMPLER_SEED = [0]    # key of the on-device Philox sampler (the reference seeds nothing: np.random global state)


def set_sampler_seed(seed):
    """Seed of the on-device Bernoulli / flipout sampler used by every engine created afterwards (and by the single-turn
    module forwards).  Data-parallel ranks use the SAME seed: their draws differ through the global row index."""
    _SAMPLER_SEED[0] = int(seed) & 0xFFFFFFFFFFFFFFFF


def _flag(name, default=None):
    try:
        return getattr(FLAGS, name)
    except Exception:
        return default


# ------------------------------------------------------------------------------------------------------------
# flags (names and defaults: model.py:1639-1741)
# ------------------------------------------------------------------------------------------------------------
def flags():
    D = gflags
    D.DEFINE_string("branch", None, ""); D.DEFINE_string("sha", None, ""); D.DEFINE_boolean("debug", False, "")
    D.DEFINE_integer("save_after", 1000, ""); D.DEFINE_integer("save_interval", 100, "")
    D.DEFINE_string("checkpoint", None, ""); D.DEFINE_string("conf_mat", None, ""); D.DEFINE_string("log_path", "./logs", "")
    D.DEFINE_string("log_file", None, ""); D.DEFINE_string("eval_csv_file", None, ""); D.DEFINE_string("json_file", None, "")
    D.DEFINE_string("log_load", None, ""); D.DEFINE_boolean("eval_only", False, "")
    D.DEFINE_boolean("binary_only", False, ""); D.DEFINE_string("binary_output", None, "")
    D.DEFINE_boolean("cuda", False, "")
    D.DEFINE_string("env", "main", ""); D.DEFINE_boolean("visdom", False, ""); D.DEFINE_boolean("use_alpha", False, "")
    D.DEFINE_string("experiment_name", None, ""); D.DEFINE_integer("log_interval", 50, ""); D.DEFINE_integer("log_dev", 1000, "")
    D.DEFINE_enum("wv_type", "glove.6B", ["fake", "glove.6B", "none"], ""); D.DEFINE_integer("wv_dim", 100, "")
    D.DEFINE_string("descr_train", "descriptions.csv", ""); D.DEFINE_string("descr_dev", "descriptions.csv", "")
    D.DEFINE_string("train_file", "train.hdf5", ""); D.DEFINE_string("dev_file", "dev.hdf5", "")
    D.DEFINE_enum("images", "mammal", ["cifar", "mammal"], "")
    D.DEFINE_string("glove_path", "./glove.6B/glove.6B.100d.txt", "")
    D.DEFINE_boolean("shuffle_train", True, ""); D.DEFINE_boolean("shuffle_dev", False, "")
    D.DEFINE_enum("model_type", None, ["Fixed", "Adaptive", "FixedAttention", "AdaptiveAttention"],
                  "Preset model configurations.")
    D.DEFINE_enum("img_feat", "avgpool_512", ["layer4_2", "avgpool_512", "fc"], "Specify which layer output to use as image")
    D.DEFINE_enum("data_context", "fc", ["fc"], "Specify which layer output to use as context for attention")
    D.DEFINE_enum("sender_mix", "sum", ["sum", "prod", "mou"], "")
    D.DEFINE_integer("img_feat_dim", 4096, ""); D.DEFINE_integer("img_h_dim", 100, "")
    D.DEFINE_integer("baseline_hid_dim", 500, ""); D.DEFINE_integer("sender_out_dim", 50, "")
    D.DEFINE_integer("rec_hidden", 128, ""); D.DEFINE_integer("rec_out_dim", 1, ""); D.DEFINE_integer("rec_w_dim", 50, "")
    D.DEFINE_integer("rec_s_dim", 1, "")
    D.DEFINE_boolean("use_binary", True, "Encoding whether Sender uses binary features")
    D.DEFINE_boolean("ignore_receiver", False, "Sender ignores messages from Receiver")
    D.DEFINE_boolean("ignore_code", False, "Sender ignores messages from Receiver")
    D.DEFINE_boolean("block_y", True, "Halt gradient flow through description scores")
    D.DEFINE_float("first_rec", 0, "")
    D.DEFINE_float("flipout_rec", None, "Dropout for bit flipping"); D.DEFINE_float("flipout_sen", None, "Dropout for bit flipping")
    D.DEFINE_boolean("flipout_dev", False, "Dropout for bit flipping")
    D.DEFINE_boolean("s_prob_prod", True, "Simulate sampling during test time")
    D.DEFINE_boolean("visual_attn", False, "Sender attends over image"); D.DEFINE_integer("attn_dim", 256, "")
    D.DEFINE_boolean("attn_extra_context", False, ""); D.DEFINE_integer("attn_context_dim", 4096, "")
    D.DEFINE_boolean("desc_attn", False, "Receiver attends over text"); D.DEFINE_integer("desc_attn_dim", 64, "Receiver attends over text")
    D.DEFINE_integer("top_k_dev", 6, "Top-k error in development"); D.DEFINE_integer("top_k_train", 6, "Top-k error in training")
    D.DEFINE_enum("optim_type", "RMSprop", ["Adam", "SGD", "RMSprop"], "")
    D.DEFINE_integer("batch_size", 32, "Minibatch size for train set."); D.DEFINE_integer("batch_size_dev", 50, "Minibatch size for dev set.")
    D.DEFINE_float("learning_rate", 1e-4, "Used in optimizer."); D.DEFINE_integer("max_epoch", 500, "")
    D.DEFINE_float("entropy_s", None, ""); D.DEFINE_float("entropy_sen", None, ""); D.DEFINE_float("entropy_rec", None, "")
    D.DEFINE_integer("exchange_samples", 3, ""); D.DEFINE_integer("max_exchange", 3, ""); D.DEFINE_boolean("fixed_exchange", True, "")
    D.DEFINE_boolean("bit_flip", False, "Whether sender's messages are corrupted.")
    D.DEFINE_string("corrupt_region", None, "Comma-separated ranges of bit indexes (e.g. ``0:3,5'').")


def Fixed():
    FLAGS.img_feat = "avgpool_512"; FLAGS.img_feat_dim = 512; FLAGS.fixed_exchange = True; FLAGS.visual_attn = False


def Adaptive():
    FLAGS.img_feat = "avgpool_512"; FLAGS.img_feat_dim = 512; FLAGS.fixed_exchange = False; FLAGS.visual_attn = False


def FixedAttention():
    FLAGS.img_feat = "layer4_2"; FLAGS.img_feat_dim = 512; FLAGS.fixed_exchange = True; FLAGS.visual_attn = True
    FLAGS.attn_dim = 256; FLAGS.attn_extra_context = True; FLAGS.attn_context_dim = 1000


def AdaptiveAttention():
    FLAGS.img_feat = "layer4_2"; FLAGS.img_feat_dim = 512; FLAGS.fixed_exchange = False; FLAGS.visual_attn = True
    FLAGS.attn_dim = 256; FLAGS.attn_extra_context = True; FLAGS.attn_context_dim = 1000


_PRESETS = dict(Fixed=Fixed, Adaptive=Adaptive, FixedAttention=FixedAttention, AdaptiveAttention=AdaptiveAttention)


def default_flags(argv=None):
    """model.py:1744-1810 (the parts that shape the path; path/log-name derivation kept)."""
    import json
    import time
    argv = sys.argv if argv is None else argv
    if FLAGS.log_load:
        log_flags = json.loads(open(FLAGS.log_load).read())
        for k in log_flags.keys():
            if k in FLAGS.flag_values_dict().keys():
                setattr(FLAGS, k, log_flags[k])
        FLAGS(argv)
    if FLAGS.model_type:
        _PRESETS[FLAGS.model_type]()
        FLAGS(argv)
    assert FLAGS.sender_out_dim == FLAGS.rec_w_dim, \
        "Both sender and receiver should communicate with same dim vectors for now."
    if not FLAGS.use_binary:
        FLAGS.exchange_samples = 0
    if not FLAGS.experiment_name:
        FLAGS.experiment_name = "{}-so_{}-wv_{}-bs_{}-{}".format(FLAGS.images, FLAGS.sender_out_dim, FLAGS.wv_dim,
                                                                FLAGS.batch_size, str(int(time.time())))
    for attr, suffix in (("conf_mat", ".conf_mat.txt"), ("log_file", ".log"), ("eval_csv_file", ".eval.csv"),
                         ("json_file", ".json"), ("checkpoint", ".pt"), ("binary_output", ".bv.hdf5")):
        if not getattr(FLAGS, attr):
            setattr(FLAGS, attr, os.path.join(FLAGS.log_path, FLAGS.experiment_name + suffix))
    if not torch.cuda.is_available():
        FLAGS.cuda = False
    if FLAGS.debug:
        np.seterr(all="raise")
    FLAGS.glove_path = os.path.expanduser(FLAGS.glove_path)


# ------------------------------------------------------------------------------------------------------------
# modules (parameters and names as the reference; forward math lives in the CUDA library)
# ------------------------------------------------------------------------------------------------------------
def xavier_normal(tensor, gain=1):
    """misc.py:367-385."""
    fan_out, fan_in = tensor.shape[0], tensor.shape[1]
    std = gain * math.sqrt(2.0 / (fan_in + fan_out))
    with torch.no_grad():
        return tensor.normal_(0, std)


def _unsupported(what):
    raise NotImplementedError("%s is outside the fused B200 path (SURVEY.md §8f); the flag is kept for CLI "
                              "compatibility only" % what)


class Sender(nn.Module):
    """Agent 1 (model.py:49-238).  Stateless; exposes `h_x` after an exchange (model.py:195,832)."""

    def __init__(self, feature_type, feat_dim, h_dim, w_dim, bin_dim_out, use_binary, use_attn=False, attn_dim=0,
                 attn_extra_context=False, attn_context_dim=0):
        super(Sender, self).__init__()
        self.feature_type, self.feat_dim, self.h_dim, self.w_dim = feature_type, feat_dim, h_dim, w_dim
        self.bin_dim_out, self.use_binary, self.use_attn = bin_dim_out, use_binary, use_attn
        self.attn_dim, self.attn_extra_context, self.attn_context_dim = attn_dim, attn_extra_context, attn_context_dim
        if use_attn:
            _unsupported("-visual_attn (Sender visual attention, model.py:80-86,114-191)")
        self.image_layer = nn.Linear(self.feat_dim, self.h_dim)
        self.code_layer = nn.Linear(self.w_dim, self.h_dim)
        self.code_bias = Parameter(torch.Tensor(self.bin_dim_out))
        if _flag("sender_mix", "sum") == "mou":           # model.py:71-76: binary_layer reads [h_x ; h_w ; h_x - h_w ; h_x * h_w]
            self.binary_layer = nn.Linear(self.h_dim * 4, self.bin_dim_out)
            if _flag("ignore_code", False):
                self.code_bias_mou = Parameter(torch.Tensor(self.bin_dim_out))
        else:
            self.binary_layer = nn.Linear(self.h_dim, self.bin_dim_out)
        self.reset_parameters()
        self.reset_state()

    def reset_parameters(self):   # model.py:90-97
        for m in self.modules():
            if isinstance(m, nn.Linear):
                xavier_normal(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()
        self.code_bias.data.normal_()
        if hasattr(self, "code_bias_mou"):
            # the reference leaves this parameter as torch.Tensor(n) allocated it (uninitialised memory, model.py:73-74, 90-97);
            # N(0, 1) like code_bias keeps runs reproducible
            self.code_bias_mou.data.normal_()

    def reset_state(self):        # model.py:99-112
        self.h_x = None
        self.attn_scores = []

    def forward(self, x, w, g, t, uniforms=None):
        """One sender turn (model.py:144-238): (message, probs-or-None); sets `self.h_x` (model.py:195).  Runs the
        stand-alone single-turn kernel (`mmg_sender_forward`); the result carries no autograd history (training goes
        through `exchange()` / `train_step()`).  `g` (visual-attention context) must be None on this path.
        `uniforms` (optional, (B,M) float64 [, flipout (B,M)]) replaces the on-device sampler by injected draws."""
        return _single_sender_step(self, x, w, g, t, uniforms)


class Receiver(nn.Module):
    """Agent 2 (model.py:241-477).  Stateful: `h_z`, `s_prob_prod`, `h_w`."""

    def __init__(self, z_dim, desc_dim, hid_dim, out_dim, w_dim, s_dim, use_binary):
        super(Receiver, self).__init__()
        self.z_dim, self.desc_dim, self.hid_dim, self.out_dim = z_dim, desc_dim, hid_dim, out_dim
        self.w_dim, self.s_dim, self.use_binary = w_dim, s_dim, use_binary
        if out_dim != 1 or s_dim != 1:
            _unsupported("rec_out_dim/rec_s_dim != 1")
        self.rnn = nn.GRUCell(self.z_dim, self.hid_dim)
        self.w_h = nn.Linear(self.hid_dim, self.hid_dim, bias=True)
        self.w_d = nn.Linear(self.desc_dim, self.hid_dim, bias=False)
        self.w = nn.Linear(self.hid_dim, self.w_dim)
        self.y1 = nn.Linear(self.hid_dim + self.desc_dim, self.hid_dim)
        self.y2 = nn.Linear(self.hid_dim, self.out_dim)
        self.s = nn.Linear(self.hid_dim, self.s_dim)
        self.desc_attn = bool(_flag("desc_attn", False))
        if self.desc_attn:                                          # model.py:267-271
            self.attn_dim = int(_flag("desc_attn_dim", 64))
            self.d_d = nn.Linear(self.desc_dim, self.attn_dim)
            self.d_h = nn.Linear(self.hid_dim, self.attn_dim)
            self.d_attn = nn.Linear(self.attn_dim, 1)
        self.reset_parameters()
        self.reset_state()

    def reset_parameters(self):   # model.py:275-288
        for m in self.modules():
            if isinstance(m, nn.Linear):
                xavier_normal(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.GRUCell):
                for mm in m.parameters():
                    if mm.data.ndimension() == 2:
                        xavier_normal(mm.data)
                    elif mm.data.ndimension() == 1:
                        mm.data.zero_()

    def reset_state(self):        # model.py:290-298
        self.h_z = None
        self.s_prob_prod = None
        self.h_w = None

    def initial_state(self, batch_size):
        return torch.zeros(batch_size, self.hid_dim, device=self.rnn.weight_ih.device)

    def forward(self, z, desc, desc_set=None, desc_set_lens=None, uniforms=None):
        """One receiver turn (model.py:303-477): ((s, s_prob), (w, w_probs), y); updates `h_z`, `h_w` and (eval mode)
        `s_prob_prod` like the reference module.  Stand-alone single-turn kernel (`mmg_receiver_forward`), no autograd
        history.  `uniforms` (optional): (u_stop (B,1), u_rec (B,M) [, flipout (B,M)]) float64 injected draws."""
        return _single_receiver_step(self, z, desc, desc_set, desc_set_lens, uniforms)


class Baseline(nn.Module):
    """REINFORCE control variate (model.py:480-516); torch-default init (no reset_parameters in the reference)."""

    def __init__(self, hid_dim, x_dim, binary_dim, inp_dim):
        super(Baseline, self).__init__()
        self.x_dim, self.binary_dim, self.inp_dim, self.hid_dim = x_dim, binary_dim, inp_dim, hid_dim
        self.linear1 = nn.Linear(x_dim + self.binary_dim + self.inp_dim, self.hid_dim)
        self.linear2 = nn.Linear(self.hid_dim, 1)

    def forward(self, x, binary, inp):
        """model.py:496-516: linear2(relu(linear1(cat(x, binary, inp)))) for whichever pieces are not None.  Stand-alone
        kernel (`mmg_baseline_forward`), no autograd history (inside `exchange()` the baselines are fused GEMM tiles)."""
        return _single_baseline_step(self, x, binary, inp)


# ------------------------------------------------------------------------------------------------------------
# binding modules <-> engine
# ------------------------------------------------------------------------------------------------------------
class _Binding(object):
    """One engine per (modules, batch, classes, flags); module parameters become views of the flat buffer."""

    def __init__(self, mods, B, D, device, batch_global=None, n_words=0, batch_offset=0):
        s, r = mods["sender"], mods["receiver"]
        cfg = _engine.make_config(batch_offset=batch_offset, 
            batch=B, n_classes=D, img_feat_dim=s.feat_dim, img_h_dim=s.h_dim, baseline_hid_dim=mods["baseline_sen"].hid_dim
            if mods.get("baseline_sen") is not None else int(_flag("baseline_hid_dim", 500)),
            sender_out_dim=s.bin_dim_out, rec_hidden=r.hid_dim, rec_w_dim=r.w_dim, wv_dim=r.desc_dim,
            max_exchange=int(_flag("max_exchange", 3)), fixed_exchange=bool(_flag("fixed_exchange", True)),
            use_binary=bool(s.use_binary), entropy_s=_flag("entropy_s"), entropy_sen=_flag("entropy_sen"),
            entropy_rec=_flag("entropy_rec"), first_rec=float(_flag("first_rec", 0) or 0),
            s_prob_prod=bool(_flag("s_prob_prod", True)), learning_rate=float(_flag("learning_rate", 1e-4)),
            optim_type=_flag("optim_type", "RMSprop"), ignore_receiver=bool(_flag("ignore_receiver", False)),
            batch_global=batch_global, flipout_sen=_flag("flipout_sen"), flipout_rec=_flag("flipout_rec"),
            flipout_dev=bool(_flag("flipout_dev", False)), sender_mix=_flag("sender_mix", "sum"),
            ignore_code=bool(_flag("ignore_code", False)), desc_attn=getattr(r, "desc_attn", False),
            desc_attn_dim=getattr(r, "attn_dim", 0), n_words=n_words)
        self.engine = _engine.GameEngine(cfg, device=device, lib=_LIB_OVERRIDE, seed=_SAMPLER_SEED[0])
        self.words_key = None
        self._mods = {a: (None if m is None else weakref.ref(m)) for a, m in mods.items()}   # engines must not keep modules alive
        self.key = None
        self.rebind()

    @property
    def mods(self):
        return {a: (None if r is None else r()) for a, r in self._mods.items()}

    def rebind(self):
        views = self.engine.named_views()
        gviews = self.engine.named_views(self.engine.grads)
        for agent, mod in self.mods.items():
            if mod is None:
                continue
            for key, p in mod.named_parameters():
                v = views[agent][key]
                if p.data_ptr() != v.data_ptr():
                    with torch.no_grad():
                        v.copy_(p.data.to(v.device).reshape(v.shape))
                    p.data = v
        self.gviews = gviews

    def publish_grads(self):
        for agent, mod in self.mods.items():
            if mod is None:
                continue
            for key, p in mod.named_parameters():
                p.grad = self.gviews[agent][key]


_BINDINGS = {}
_LAST_BINDING = {}     # module set -> the binding that trained last: owner of the freshest optimizer state


def _mods_key(sender, receiver, baseline_sen, baseline_rec):
    return (id(sender), id(receiver), id(baseline_sen), id(baseline_rec))


def _drop_bindings(mkey):
    for k in [k for k in _BINDINGS if k[:4] == mkey]:
        del _BINDINGS[k]
    _LAST_BINDING.pop(mkey, None)


def _adopt_state(dst, src):
    """The fused optimizer state (RMSprop / Adam moments, step counts) does not depend on the batch size: an engine created
    for another batch size or flag value of the SAME modules continues from the state of the engine that trained last
    instead of starting from zero."""
    a, b = dst.engine, src.engine
    if a is b or int(a.layout.total) != int(b.layout.total) or int(a.cfg.optim_type) != int(b.cfg.optim_type):
        return
    with torch.no_grad():
        a.state1.copy_(b.state1)
        if a.state2 is not None and b.state2 is not None:
            a.state2.copy_(b.state2)
        a.ws("opt_counters", (4,), torch.int64).copy_(b.ws("opt_counters", (4,), torch.int64))
    a.step = b.step


def _binding_for(sender, receiver, baseline_sen, baseline_rec, B, D, device, batch_global=None, exchange_args=None,
                 batch_offset=0):
    """One engine per shape/flag combination.  With -desc_attn the word set of the class descriptions
    (exchange_args["desc_set"], ["desc_set_lens"], model.py:765-766) belongs to the binding: its size is part of the
    configuration and the words are uploaded once per distinct set (train / dev sets differ)."""
    n_words = 0
    if getattr(receiver, "desc_attn", False):
        desc_set = (exchange_args or {}).get("desc_set")
        lens = (exchange_args or {}).get("desc_set_lens")
        if desc_set is None or lens is None:
            raise ValueError("-desc_attn needs exchange_args['desc_set'] and ['desc_set_lens']")
        n_words = int(desc_set.shape[0])
    b = _binding_lookup(sender, receiver, baseline_sen, baseline_rec, B, D, device, batch_global, n_words, batch_offset)
    if n_words:
        wk = (id(desc_set), int(desc_set.data_ptr()) if torch.is_tensor(desc_set) else 0, tuple(int(v) for v in lens))
        if b.words_key != wk:
            b.engine.set_desc_set(desc_set, lens)
            b.words_key = wk
    return b


def _binding_lookup(sender, receiver, baseline_sen, baseline_rec, B, D, device, batch_global, n_words, batch_offset=0):
    mkey = _mods_key(sender, receiver, baseline_sen, baseline_rec)
    # every flag that reaches mmg_config is part of the key: a changed flag never reuses a stale configuration
    key = mkey + (int(B), int(D), str(device),
           int(_flag("max_exchange", 3)), bool(_flag("fixed_exchange", True)), _flag("entropy_s"), _flag("entropy_sen"),
           _flag("entropy_rec"), _flag("optim_type", "RMSprop"), float(_flag("learning_rate", 1e-4)), batch_global,
           _flag("flipout_sen"), _flag("flipout_rec"), bool(_flag("flipout_dev", False)), _flag("sender_mix", "sum"),
           bool(_flag("ignore_code", False)), n_words, float(_flag("first_rec", 0) or 0), bool(_flag("ignore_receiver", False)),
           bool(_flag("s_prob_prod", True)), int(_flag("baseline_hid_dim", 500)), bool(sender.use_binary), int(batch_offset),
           _SAMPLER_SEED[0])
    b = _BINDINGS.get(key)
    if b is None:
        mods = dict(receiver=receiver, sender=sender, baseline_rec=baseline_rec, baseline_sen=baseline_sen)
        b = _Binding(mods, B, D, device, batch_global, n_words, batch_offset)
        if not any(k[:4] == mkey for k in _BINDINGS):
            weakref.finalize(sender, _drop_bindings, mkey)      # engines die with their modules
        _BINDINGS[key] = b
    else:
        b.rebind()
    last = _LAST_BINDING.get(mkey)
    if last is not None and last is not b:
        _adopt_state(b, last)
    return b


def build_mask(region_str, size):
    """misc.py:388-402."""
    mask = torch.zeros(size)
    for r in region_str.split(","):
        r = r.split(":")
        idx = [int(r[0])] if len(r) == 1 else list(range(int(r[0]), int(r[1])))
        mask[idx] = 1
    return mask


_AGENT_OUTPUTS = {"receiver": ("rec_probs", "stop_prob", "y"), "sender": ("sen_probs",), "baseline_rec": ("br",),
                  "baseline_sen": ("bs",)}


class _AgentFn(torch.autograd.Function):
    """Autograd bridge, one node per module: the outputs of the (already executed) fused conversation that carry
    gradient to THIS module's parameters.  The reference's four losses are disjoint graphs (every cross-agent tensor is
    cut with `.data`, model.py:807-843), so each of its `loss.backward()` calls (model.py:1309,1316,1322,1328) reaches
    exactly one of these nodes and runs the fused backward kernels for that module."""

    @staticmethod
    def forward(ctx, binding, agent, inp, *params):
        o = binding.engine.outputs()
        ctx.binding, ctx.agent, ctx.inp = binding, agent, inp
        return tuple(o[name].clone() for name in _AGENT_OUTPUTS[agent])

    @staticmethod
    def backward(ctx, *gouts):
        b = ctx.binding
        e = b.engine
        d = e.dims
        T, B, M, D = d["T"], d["B"], d["M"], d["D"]
        g = dict(zip(_AGENT_OUTPUTS[ctx.agent], gouts))
        z = lambda name, shape: torch.zeros(shape, device=e.device) if g.get(name) is None else g[name]
        e.ws("g_sen_probs", (T, B, M)).copy_(z("sen_probs", (T, B, M)))
        e.ws("g_rec_probs", (T, B, M)).copy_(z("rec_probs", (T, B, M)))
        e.ws("g_stop_prob", (T, B)).copy_(z("stop_prob", (T, B, 1)).reshape(T, B))
        e.ws("g_bs", (T, B)).copy_(z("bs", (T, B, 1)).reshape(T, B))
        e.ws("g_br", (T, B)).copy_(z("br", (T, B, 1)).reshape(T, B))
        gy = z("y", (T, B, D))
        nz = (gy != 0).any(dim=2)                                  # (T, B): steps whose scores received gradient
        if bool((nz.sum(0) > 1).any()):
            _unsupported("gradients into the class scores of more than one step per example")
        ystep = torch.where(nz.any(0), nz.float().argmax(0), torch.full((B,), T - 1, device=e.device)).to(torch.int32)
        e.ws("ystep", (B,), torch.int32).copy_(ystep)
        e.ws("g_outp", (B, D)).copy_(gy[ystep.long(), torch.arange(B, device=e.device)])
        e._inp = ctx.inp
        e.backward()
        gv = b.gviews
        grads = [gv[ctx.agent][key].clone() for key in e.param_keys(ctx.agent)]
        return (None, None, None) + tuple(grads)


def _device_of(module):
    return next(module.parameters()).device


def exchange(sender, receiver, baseline_sen, baseline_rec, exchange_args):
    """Batched conversation (model.py:725-876).  Same arguments and return structure as the reference:
    s = (stop_mask[T'+1], stop_feat[T'], stop_prob[T']), sen_w = (feats, probs), rec_w = (feats, probs), y, bs, br.
    Optional extra key exchange_args["uniforms"] = (u_sen (T,B,M), u_stop (T,B), u_rec (T,B,M)) float64 replaces the
    on-device sampler by injected draws in the reference's order (parity tests)."""
    data = exchange_args["data"]
    target = exchange_args.get("target")
    desc = exchange_args["desc"]
    train = exchange_args["train"]
    break_early = exchange_args.get("break_early", False)
    corrupt = exchange_args.get("corrupt", False)
    corrupt_region = exchange_args.get("corrupt_region", None)
    if exchange_args.get("data_context") is not None:
        _unsupported("data_context (visual attention)")
    dev = _device_of(receiver)
    B, D = data.shape[0], desc.shape[0]
    binding = _binding_for(sender, receiver, baseline_sen, baseline_rec, B, D, dev, exchange_args=exchange_args)
    e = binding.engine
    T = e.dims["T"]
    if train:
        sender.train(); receiver.train(); baseline_sen.train(); baseline_rec.train()    # model.py:789-793
    else:
        sender.eval(); receiver.eval()
    sender.reset_state(); receiver.reset_state()
    mask = build_mask(corrupt_region, sender.w_dim).to(dev) if corrupt else None
    binary = bool(sender.use_binary)
    if train:
        with torch.no_grad():
            e.forward(data.detach(), desc.detach(), target, train=True, uniforms=exchange_args.get("uniforms"))
        outs = {}
        for agent in capi.SEGMENTS:
            named = dict(binding.mods[agent].named_parameters())
            params = [named[key] for key in e.param_keys(agent)]
            res = _AgentFn.apply(binding, agent, e._inp, *params)
            outs.update(zip(_AGENT_OUTPUTS[agent], res))
        sen_p, rec_p, stop_p, y_all, bs_all, br_all = (outs["sen_probs"], outs["rec_probs"], outs["stop_prob"], outs["y"],
                                                       outs["bs"], outs["br"])
        o = e.outputs()
    else:
        with torch.no_grad():
            e.forward(data, desc, target, train=False, corrupt_mask=mask)
        o = e.outputs()
        sen_p, rec_p, stop_p, y_all = o["sen_probs"].clone(), o["rec_probs"].clone(), o["stop_prob"].clone(), o["y"].clone()
    masks = o["stop_mask"].clone()                       # (T+1, B, 1) uint8, raw chain
    steps = T
    if break_early:                                       # model.py:866: stop once nobody is active
        alive = masks[1:].reshape(T, -1).sum(1)
        dead = (alive == 0).nonzero()
        if dead.numel() > 0:
            steps = int(dead[0]) + 1
    stop_mask = [masks[t] for t in range(steps + 1)]
    stop_mask[-1] = torch.zeros_like(stop_mask[-1])       # model.py:870
    sen_feats_all, rec_feats_all, stop_feat_all = o["sen_feats"].clone(), o["rec_feats"].clone(), o["stop_feat"].clone()
    stop_feat = [stop_feat_all[t] for t in range(steps)]
    stop_prob = [stop_p[t] for t in range(steps)]
    sen_feats = [sen_feats_all[t] for t in range(steps)]
    rec_feats = [rec_feats_all[t] for t in range(steps)]
    sen_probs = [sen_p[t] if binary else None for t in range(steps)]
    rec_probs = [rec_p[t] if binary else None for t in range(steps)]
    y = [y_all[t] for t in range(steps)]
    bs = [bs_all[t] for t in range(steps)] if train else []
    br = [br_all[t] for t in range(steps)] if train else []
    sender.h_x = o["h_x"]
    receiver.h_z = o["h_z"][steps - 1]
    receiver.h_w = o["h_w"][steps - 1]
    return (stop_mask, stop_feat, stop_prob), (sen_feats, sen_probs), (rec_feats, rec_probs), y, bs, br


def eval_dev(dev_file, batch_size, epoch, shuffle, cuda, top_k, sender, receiver, desc_dict, map_labels, file_name,
             callback=None):
    """Development accuracy and conversation statistics (model.py:580-722), same signature and return value
    (accuracy, extra).  The conversations run in eval mode through the fused kernels; top-k membership, conversation
    lengths, Hamming distances between consecutive messages and the confusion matrix are accumulated ON THE DEVICE and
    read back once at the end (the reference copies every batch to the host and argsorts with NumPy).

    `dev_file` is either an iterable of batch dicts in the layout `misc.load_hdf5` yields (keys "target" and
    FLAGS.img_feat, model.py:614-615) or an HDF5 path (needs h5py, which this package does not depend on)."""
    if isinstance(dev_file, str):
        try:
            import h5py  # noqa: F401
        except ImportError:
            raise ImportError("eval_dev(dev_file=<path>) needs h5py (misc.load_hdf5, misc.py:257-302); pass an iterable "
                              "of batch dicts instead")
        _unsupported("reading HDF5 feature files (data loading is outside the fused path; pass batch dicts)")
    dev = _device_of(receiver)
    desc = desc_dict["desc"].to(dev)
    n_classes = desc.shape[0]
    feat_key = _flag("img_feat", "avgpool_512")
    fixed = bool(_flag("fixed_exchange", True))
    total = 0.0
    correct = torch.zeros((), dtype=torch.float64, device=dev)
    conv_sum = torch.zeros((), dtype=torch.float64, device=dev)
    conv_sq = torch.zeros((), dtype=torch.float64, device=dev)
    conv_n = 0
    ham_sen, ham_rec = [], []
    conf = torch.zeros(n_classes, n_classes, dtype=torch.int64, device=dev)
    for batch in dev_file:
        target = batch["target"].to(dev)
        data = batch[feat_key].to(dev)
        bsz = target.size(0)
        exchange_args = dict(data=data, target=target, desc=desc, desc_set=desc_dict.get("desc_set"),
                             desc_set_lens=desc_dict.get("desc_set_lens"), train=False, break_early=not fixed,
                             corrupt=bool(_flag("bit_flip", False)), corrupt_region=_flag("corrupt_region"))
        s, sen_w, rec_w, y, bs, br = exchange(sender, receiver, None, None, exchange_args)
        s_masks, s_feats, s_probs = s
        sen_feats, sen_probs = sen_w
        rec_feats, rec_probs = rec_w
        y_masks = None if fixed else [torch.min(1 - m1, m2) for m1, m2 in zip(s_masks[1:], s_masks[:-1])]   # 648-652
        outp, _ = get_rec_outp(y, y_masks)
        dist = F.log_softmax(outp, dim=1)
        k = min(int(top_k), n_classes)
        top_k_ind = dist.topk(k, dim=1).indices                                   # model.py:658-660
        correct += (top_k_ind == target.view(-1, 1)).sum()
        total += float(batch_size)                                                # model.py:668 (nominal batch size)
        argmax = dist.argmax(dim=1)
        conf += torch.bincount(target * n_classes + argmax, minlength=n_classes * n_classes).view(n_classes, n_classes)
        lengths = torch.cat(s_feats, 1).float().sum(1).double()                   # model.py:672-673
        conv_sum += lengths.sum(); conv_sq += (lengths * lengths).sum(); conv_n += bsz
        for feats, acc in ((sen_feats, ham_sen), (rec_feats, ham_rec)):            # model.py:676-690
            msgs = torch.stack(feats, 0)
            prev = torch.cat([torch.zeros_like(msgs[:1]), msgs[:-1]], 0)
            acc.append((msgs - prev).abs().sum(2).mean(1).sum() / float(len(feats)))
        if callback is not None:
            callback(sender, receiver, batch, dict(s_masks=s_masks, s_feats=s_feats, s_probs=s_probs, sen_feats=sen_feats,
                                                   sen_probs=sen_probs, rec_feats=rec_feats, rec_probs=rec_probs, y=y))
    mean = float(conv_sum) / max(conv_n, 1)
    extra = dict()
    extra["conversation_lengths_mean"] = mean
    extra["conversation_lengths_std"] = math.sqrt(max(float(conv_sq) / max(conv_n, 1) - mean * mean, 0.0))
    extra["hamming_sen_mean"] = float(torch.stack(ham_sen).mean()) if ham_sen else float("nan")
    extra["hamming_rec_mean"] = float(torch.stack(ham_rec).mean()) if ham_rec else float("nan")
    extra["confusion_matrix"] = conf.cpu().numpy()
    conf_path = _flag("conf_mat")
    if conf_path:
        try:
            np.savetxt(conf_path, extra["confusion_matrix"], delimiter=",", fmt="%d")   # model.py:709-710
        except (IOError, OSError):
            pass
    return float(correct) / max(total, 1.0), extra


# ------------------------------------------------------------------------------------------------------------
# checkpoints (misc.py:42-92): same dictionary layout {data, optimizers, models}, tensors always saved on the CPU
# ------------------------------------------------------------------------------------------------------------
def recursively_set_device(inp, gpu):
    """misc.py:42-55."""
    if hasattr(inp, "keys"):
        for k in inp.keys():
            inp[k] = recursively_set_device(inp[k], gpu)
    elif isinstance(inp, list):
        return [recursively_set_device(ii, gpu) for ii in inp]
    elif isinstance(inp, tuple):
        return tuple(recursively_set_device(ii, gpu) for ii in inp)
    elif hasattr(inp, "cpu"):
        inp = inp.cuda() if gpu >= 0 else inp.cpu()
    return inp


def torch_save(filename, data, models_dict, optimizers_dict, gpu):
    """misc.py:58-75.  `optimizers_dict` values are torch.optim objects or `FusedOptimizer` views."""
    models_to_save = {k: recursively_set_device({kk: vv.detach().clone() for kk, vv in v.state_dict().items()}, gpu=-1)
                      for k, v in models_dict.items()}
    optimizers_to_save = {k: recursively_set_device(v.state_dict(), gpu=-1) for k, v in optimizers_dict.items()}
    torch.save({"data": data, "optimizers": optimizers_to_save, "models": models_to_save}, filename)


def torch_load(filename, models_dict, optimizers_dict):
    """misc.py:78-92.  Module parameters are views of the engine's flat buffer: loading copies INTO them."""
    filename = os.path.expanduser(filename)
    if not os.path.exists(filename):
        raise Exception("File does not exist: " + filename)
    checkpoint = torch.load(filename, weights_only=False)
    for k, v in models_dict.items():
        with torch.no_grad():
            for key, p in v.named_parameters():
                p.copy_(checkpoint["models"][k][key].to(p.device))
    for k, v in optimizers_dict.items():
        v.load_state_dict(checkpoint["optimizers"][k])
    return checkpoint["data"]


class FusedOptimizer(object):
    """torch.optim-shaped view (state_dict / load_state_dict) of ONE module's slice of the fused optimizer state that
    `train_step()` keeps in the engine, so the reference's checkpoint code (model.py:1139-1156,1569-1584) keeps working
    with `optimizers_dict = dict(optimizer_rec=FusedOptimizer(engine, "receiver"), ...)`.  The layout follows
    torch.optim.RMSprop / Adam: state[i] for the i-th parameter of the module in `parameters()` order."""

    def __init__(self, engine, agent):
        self.engine, self.agent = engine, agent
        self.keys = engine.param_keys(agent)
        # parameters() order = state_dict order (direct parameters first), which is how PARAM_NAMES is laid out
        self.optim = {v: k for k, v in capi.OPTIM.items()}[int(engine.cfg.optim_type)]

    def _views(self):
        e = self.engine
        s1 = e.named_views(e.state1)[self.agent]
        s2 = e.named_views(e.state2)[self.agent] if e.state2 is not None else None
        return s1, s2

    def state_dict(self):
        e = self.engine
        s1, s2 = self._views()
        state = {}
        head_steps = float(e.ws("opt_counters", (4,), torch.int64)[0]) if self.agent == "receiver" else 0.0
        for i, key in enumerate(self.keys):
            in_head = self.agent == "receiver" and key.split(".")[0] in ("w_h", "w_d", "w")
            st = {"step": torch.tensor(head_steps if in_head else float(e.step))}
            if self.optim == "RMSprop":
                st["square_avg"] = s1[key].detach().clone()
            elif self.optim == "Adam":
                st["exp_avg"] = s2[key].detach().clone()
                st["exp_avg_sq"] = s1[key].detach().clone()
            state[i] = st
        group = dict(lr=float(e.cfg.learning_rate), params=list(range(len(self.keys))))
        if self.optim == "RMSprop":
            group.update(alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False)
        elif self.optim == "Adam":
            group.update(betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False)
        return {"state": state, "param_groups": [group], "fused": {"optim": self.optim, "agent": self.agent}}

    def load_state_dict(self, sd):
        e = self.engine
        s1, s2 = self._views()
        with torch.no_grad():
            for i, key in enumerate(self.keys):
                st = sd["state"].get(i, sd["state"].get(str(i)))
                if st is None:
                    continue
                if self.optim == "RMSprop":
                    s1[key].copy_(st["square_avg"].to(s1[key].device).reshape(s1[key].shape))
                elif self.optim == "Adam":
                    s2[key].copy_(st["exp_avg"].to(s2[key].device).reshape(s2[key].shape))
                    s1[key].copy_(st["exp_avg_sq"].to(s1[key].device).reshape(s1[key].shape))
                if "step" in st:
                    if self.agent == "receiver" and key.split(".")[0] in ("w_h", "w_d", "w"):
                        e.ws("opt_counters", (4,), torch.int64)[0] = int(float(st["step"]))
                    else:
                        e.step = int(float(st["step"]))


def train_step(sender, receiver, baseline_sen, baseline_rec, exchange_args, group=None):
    """The whole iteration of run() (model.py:1240-1339) fused on the device: conversation, the five losses, backward,
    per-module clip_grad_norm(1.) and the optimizer step.  Returns the engine (losses()/outputs() read results).
    With `group` (torch.distributed, NCCL) the batch in `exchange_args` is this rank's shard of a global batch of
    world_size * B rows: batch statistics and the flat gradient buffer are all-reduced."""
    data, target, desc = exchange_args["data"], exchange_args["target"], exchange_args["desc"]
    dev = _device_of(receiver)
    world = 1
    if group is not None:
        import torch.distributed as dist
        world = dist.get_world_size(group)
    rank = dist.get_rank(group) if world > 1 else 0
    binding = _binding_for(sender, receiver, baseline_sen, baseline_rec, data.shape[0], desc.shape[0], dev,
                           batch_global=data.shape[0] * world if world > 1 else None, exchange_args=exchange_args,
                           batch_offset=rank * data.shape[0])
    _LAST_BINDING[_mods_key(sender, receiver, baseline_sen, baseline_rec)] = binding
    e = binding.engine
    top_k = min(int(_flag("top_k_train", 6)), desc.shape[0])
    if world > 1:
        e.train_step_dp(data, desc, target, group=group, uniforms=exchange_args.get("uniforms"), top_k=top_k)
    else:
        e.train_step(data, desc, target, uniforms=exchange_args.get("uniforms"), top_k=top_k)
    binding.publish_grads()
    return e


# ------------------------------------------------------------------------------------------------------------
# single-turn module forwards (no autograd): one-step conversations through the same kernels
# ------------------------------------------------------------------------------------------------------------
class _ModuleFlat(object):
    """Flat parameter buffer (the C-ABI's layout) of ONE stand-alone module: the single-turn entry points read the
    module's tensors from it; the other modules' slots stay zero and are never read."""

    def __init__(self, module, agent, dims):
        self.lib = _LIB_OVERRIDE if _LIB_OVERRIDE is not None else capi.load()
        self.agent = agent
        self.cfg = _engine.make_config(batch=1, **dims)
        self.layout = capi.ParamLayout()
        self.lib.call("mmg_param_layout_get", C.byref(self.cfg), C.byref(self.layout))
        self.device = _device_of(module)
        self.flat = torch.zeros(int(self.layout.total), dtype=torch.float32, device=self.device)
        self.slots = {}
        for idx, (a, key) in enumerate(capi.PARAM_NAMES):
            n = int(self.layout.rows[idx]) * int(self.layout.cols[idx])
            if a == agent and n:
                off = int(self.layout.offset[idx])
                self.slots[key] = self.flat[off:off + n]
        self.counter = 0

    def refresh(self, module):
        with torch.no_grad():
            for key, p in module.named_parameters():
                self.slots[key].copy_(p.detach().reshape(-1))

    def stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)


def _flat_for(module, agent, dims):
    key = tuple(sorted(dims.items()))
    mf = module.__dict__.get("_mmg_flat")
    if mf is None or mf[0] != key or mf[1].device != _device_of(module):
        mf = (key, _ModuleFlat(module, agent, dims))
        module.__dict__["_mmg_flat"] = mf
    mf[1].refresh(module)
    return mf[1]


def _f32(t, dev):
    return None if t is None else torch.as_tensor(t).detach().to(device=dev, dtype=torch.float32).contiguous()


def _f64(t, dev):
    return None if t is None else torch.as_tensor(t).detach().to(device=dev, dtype=torch.float64).contiguous()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _flip_flags():
    return dict(flipout_sen=_flag("flipout_sen"), flipout_rec=_flag("flipout_rec"), flipout_dev=bool(_flag("flipout_dev", False)))


def _single_sender_step(sender, x, w, g, t, uniforms=None):
    if g is not None:
        _unsupported("data_context (visual attention)")
    mf = _flat_for(sender, "sender", dict(
        n_classes=1, img_feat_dim=sender.feat_dim, img_h_dim=sender.h_dim, sender_out_dim=sender.bin_dim_out,
        rec_w_dim=sender.bin_dim_out, rec_hidden=1, wv_dim=1, baseline_hid_dim=1, use_binary=bool(sender.use_binary),
        sender_mix=_flag("sender_mix", "sum"), ignore_code=bool(_flag("ignore_code", False)), **_flip_flags()))
    dev = mf.device
    xx = _f32(x, dev)
    B, M = xx.shape[0], sender.bin_dim_out
    ww = _f32(w, dev) if t != 0 else None
    us = [_f64(u, dev) for u in (uniforms or ())] + [None, None]
    msg = torch.empty(B, M, device=dev)
    probs = torch.empty(B, M, device=dev) if sender.use_binary else None
    h_x = torch.empty(B, sender.h_dim, device=dev)
    mf.counter += 1
    mf.lib.call("mmg_sender_forward", C.byref(mf.cfg), mf.flat.data_ptr(), B, _ptr(xx), _ptr(ww), int(t), int(sender.training),
                _ptr(us[0]), _ptr(us[1]), C.c_uint64(_SAMPLER_SEED[0]), C.c_uint64(2 * mf.counter),
                _ptr(msg), _ptr(probs), _ptr(h_x), mf.stream())
    sender.h_x = h_x
    return msg, probs


def _single_receiver_step(receiver, z, desc, desc_set=None, desc_set_lens=None, uniforms=None):
    attn = bool(getattr(receiver, "desc_attn", False))
    n_words = int(desc_set.shape[0]) if attn else 0
    if attn and (desc_set is None or desc_set_lens is None):
        raise ValueError("-desc_attn needs desc_set and desc_set_lens")
    mf = _flat_for(receiver, "receiver", dict(
        n_classes=int(desc.shape[0]), img_feat_dim=1, img_h_dim=1, sender_out_dim=receiver.w_dim, rec_w_dim=receiver.w_dim,
        rec_hidden=receiver.hid_dim, wv_dim=receiver.desc_dim, baseline_hid_dim=1, use_binary=bool(receiver.use_binary),
        s_prob_prod=bool(_flag("s_prob_prod", True)), ignore_receiver=bool(_flag("ignore_receiver", False)),
        desc_attn=attn, desc_attn_dim=getattr(receiver, "attn_dim", 0), n_words=n_words, **_flip_flags()))
    dev = mf.device
    zz, dd = _f32(z, dev), _f32(desc, dev)
    B, M, Hr, D = zz.shape[0], receiver.w_dim, receiver.hid_dim, dd.shape[0]
    first = receiver.h_z is None                                   # model.py:336-337
    h_z = torch.zeros(B, Hr, device=dev) if first else _f32(receiver.h_z, dev).clone()
    train = bool(receiver.training)
    sprod, first_prod = None, True
    if not train:
        first_prod = receiver.s_prob_prod is None                   # model.py:423
        sprod = torch.zeros(B, device=dev) if first_prod else _f32(receiver.s_prob_prod, dev).reshape(B).clone()
    ds = _f32(desc_set, dev) if attn else None
    dl = torch.as_tensor([int(v) for v in desc_set_lens], dtype=torch.int32, device=dev) if attn else None
    us = [_f64(u, dev) for u in (uniforms or ())] + [None, None, None]
    s, s_prob = torch.empty(B, 1, device=dev), torch.empty(B, 1, device=dev)
    wv = torch.empty(B, M, device=dev)
    w_probs = torch.empty(B, M, device=dev) if receiver.use_binary else None
    y, h_w = torch.empty(B, D, device=dev), torch.empty(B, Hr, device=dev)
    mf.counter += 1
    # `first` also resets the running STOP product: both states are cleared together by reset_state() (model.py:290-298)
    mf.lib.call("mmg_receiver_forward", C.byref(mf.cfg), mf.flat.data_ptr(), B, _ptr(zz), _ptr(dd), _ptr(ds), _ptr(dl),
                _ptr(h_z), _ptr(sprod), int(first and first_prod), int(train), _ptr(us[0]), _ptr(us[1]), _ptr(us[2]),
                C.c_uint64(_SAMPLER_SEED[0]), C.c_uint64(2 * mf.counter + 1), _ptr(s), _ptr(s_prob), _ptr(wv),
                _ptr(w_probs), _ptr(y), _ptr(h_w), mf.stream())
    receiver.h_z, receiver.h_w = h_z, h_w
    if not train:
        receiver.s_prob_prod = sprod.view(B, 1)
    return (s, s_prob), (wv, w_probs), y


def _single_baseline_step(baseline, x, binary, inp):
    which = 3 if baseline.x_dim > 0 else 2            # MMG_SEG_BASELINE_SEN: [h_x ; z_r]; MMG_SEG_BASELINE_REC: [z ; h_z]
    if which == 3:
        dims = dict(img_h_dim=baseline.x_dim, sender_out_dim=baseline.binary_dim, rec_w_dim=baseline.binary_dim, rec_hidden=1)
        if baseline.inp_dim:
            _unsupported("Baseline with all three inputs (the game only builds (x, binary) and (binary, inp) baselines)")
    else:
        dims = dict(img_h_dim=1, sender_out_dim=baseline.binary_dim, rec_w_dim=baseline.binary_dim, rec_hidden=baseline.inp_dim)
    mf = _flat_for(baseline, "baseline_sen" if which == 3 else "baseline_rec",
                   dict(n_classes=1, img_feat_dim=1, wv_dim=1, baseline_hid_dim=baseline.hid_dim, **dims))
    dev = mf.device
    xx, bb, ii = _f32(x, dev), _f32(binary, dev), _f32(inp, dev)
    rows = next(t for t in (xx, bb, ii) if t is not None).shape[0]
    out = torch.empty(rows, 1, device=dev)
    width = lambda t: 0 if t is None else int(t.shape[1])
    mf.lib.call("mmg_baseline_forward", C.byref(mf.cfg), mf.flat.data_ptr(), which, rows, _ptr(xx), width(xx), _ptr(bb), width(bb),
                _ptr(ii), width(ii), _ptr(out), mf.stream())
    return out


# ------------------------------------------------------------------------------------------------------------
# reference loss surface (model.py:519-577, 879-988) for callers that keep the reference's own update block
# ------------------------------------------------------------------------------------------------------------
def build_inp(binary_features, descs):
    """model.py:519-551 (kept for API compatibility; the fused path never materialises this product)."""
    if descs is None:
        return binary_features
    B, D = binary_features.size(0), descs.size(0)
    hb = binary_features.unsqueeze(1).expand(B, D, binary_features.size(1)).reshape(B * D, -1)
    dd = descs.unsqueeze(0).expand(B, D, descs.size(1)).reshape(B * D, -1)
    return torch.cat([hb, dd], 1)


def flipout(binary, p):
    """model.py:554-568."""
    mask = torch.from_numpy((np.random.rand(*binary.shape) < p).astype("float32")).to(binary.device)
    return (binary - mask).abs()


def loglikelihood(log_prob, target):
    return log_prob.gather(1, target)


def get_rec_outp(y, masks):
    """model.py:879-904."""
    def negent(yy):
        probs = F.softmax(yy, dim=1)
        return (torch.log(probs + 1e-8) * probs).sum(1).mean()
    negentropy = [negent(yy) for yy in y]
    if masks is None:
        return y[-1], negentropy
    B = y[0].size(0)
    inp = torch.stack(y, 1)
    m = torch.cat(masks, 1).bool()
    if _flag("debug", False):
        assert bool((m.sum(1) == 1).all())
    return inp[m].view(B, -1), negentropy


def calculate_loss_binary(binary_features, binary_probs, logs, baseline_scores, entropy_penalty):
    """model.py:907-927 with the torch-0.1.12 shapes made explicit ((B,1) per-example terms, no broadcasting)."""
    f = binary_features.detach()
    log_p_z = (f * torch.log(binary_probs + 1e-8) + (1 - f) * torch.log(1 - binary_probs + 1e-8)).sum(1, keepdim=True)
    weight = logs.detach() - baseline_scores.detach()
    if logs.size(0) > 1:
        weight = weight / max(1.0, float(torch.std(weight)))
    loss = torch.mean(-1 * weight * log_p_z)
    initial_negent = (torch.log(binary_probs + 1e-8) * binary_probs).sum(1).mean()
    inverse_negent = (torch.log((1. - binary_probs) + 1e-8) * (1. - binary_probs)).sum(1).mean()
    negentropy = initial_negent + inverse_negent
    if entropy_penalty is not None:
        loss = loss + entropy_penalty * negentropy
    return loss, negentropy


def multistep_loss_binary(binary_features, binary_probs, logs, baseline_scores, masks, entropy_penalty):
    """model.py:930-968."""
    if masks is not None:
        sums = [float(m.float().sum()) for m in masks]
        losses, entropies = [], []
        for feat, prob, scores, mask, ms in zip(binary_features, binary_probs, baseline_scores, masks, sums):
            if ms == 0:
                losses.append(torch.zeros((), device=logs.device))
                continue
            sel = mask.view(-1).bool()
            l, e = calculate_loss_binary(feat[sel], prob[sel], logs[sel], scores[sel], entropy_penalty)
            losses.append(l); entropies.append(e)
        loss = sum(l * ms for l, ms in zip(losses, sums)) / sum(sums)
    else:
        outp = [calculate_loss_binary(feat, prob, logs, scores, entropy_penalty)
                for feat, prob, scores in zip(binary_features, binary_probs, baseline_scores)]
        losses = [o[0] for o in outp]
        entropies = [o[1] for o in outp]
        loss = sum(losses) / len(binary_features)
    return loss, entropies


def calculate_loss_bas(baseline_scores, logs):
    return F.mse_loss(baseline_scores, logs.detach())


def multistep_loss_bas(baseline_scores, logs, masks):
    """model.py:976-988."""
    if masks is not None:
        losses, sums = [], []
        for scores, mask in zip(baseline_scores, masks):
            sel = mask.view(-1).bool()
            losses.append(calculate_loss_bas(scores[sel].view(-1, 1), logs[sel].view(-1, 1)))
            sums.append(float(mask.float().sum()))
        return sum(l * ms for l, ms in zip(losses, sums)) / sum(sums)
    losses = [calculate_loss_bas(scores, logs) for scores in baseline_scores]
    return sum(losses) / len(baseline_scores)
