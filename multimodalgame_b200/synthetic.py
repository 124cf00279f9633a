"""Synthetic workloads for the game (product side: used by bench.py, __graft_entry__.smoke() and the examples).

The reference trains on pre-extracted image features and GloVe class descriptions that are not available offline; the
BASELINE.json configurations are therefore run on synthetic data of the same shapes: x ~ N(0,1) (B,F),
desc ~ N(0,1) (D,WV) as `-wv_type fake` does (model.py:1069), targets uniform, parameters initialised as the reference's
modules initialise them (model.py:90-97, 275-288: Xavier-normal weights, zero biases, code_bias ~ N(0,1); the baselines
keep nn.Linear's default init).  Nothing here touches the oracle: the same tensors are handed to both arms of a
comparison by the caller.
"""
import math
from collections import OrderedDict

import torch

from . import engine as _engine

# flag names and defaults of the reference that shape the path (model.py:1641-1741)
FLAG_DEFAULTS = OrderedDict(
    batch_size=32, img_feat_dim=4096, img_h_dim=100, baseline_hid_dim=500, sender_out_dim=50, rec_hidden=128, rec_w_dim=50,
    wv_dim=100, n_classes=30, max_exchange=3, fixed_exchange=True, use_binary=True, entropy_s=None, entropy_sen=None,
    entropy_rec=None, first_rec=0.0, s_prob_prod=True, learning_rate=1e-4, optim_type="RMSprop", top_k_train=6,
    ignore_receiver=False, flipout_sen=None, flipout_rec=None, sender_mix="sum", ignore_code=False, desc_attn=False,
    desc_attn_dim=64)


class GameFlags(object):
    """The subset of the reference's gflags that shapes the hot path, as plain attributes."""

    def __init__(self, **kw):
        unknown = set(kw) - set(FLAG_DEFAULTS)
        if unknown:
            raise TypeError("unknown flags: %s" % sorted(unknown))
        for k, v in FLAG_DEFAULTS.items():
            setattr(self, k, kw.get(k, v))
        assert self.sender_out_dim == self.rec_w_dim, \
            "Both sender and receiver should communicate with same dim vectors for now."   # model.py:1756

    def as_dict(self):
        return OrderedDict((k, getattr(self, k)) for k in FLAG_DEFAULTS)


def config_from_flags(flags, batch=None, batch_global=None, n_words=0, batch_offset=0):
    """mmg_config for an object carrying the reference's flag names (GameFlags, gflags.FLAGS, the tests' GameConfig)."""
    g = lambda name: getattr(flags, name, FLAG_DEFAULTS[name])
    return _engine.make_config(
        batch=batch or g("batch_size"), n_classes=g("n_classes"), img_feat_dim=g("img_feat_dim"), img_h_dim=g("img_h_dim"),
        baseline_hid_dim=g("baseline_hid_dim"), sender_out_dim=g("sender_out_dim"), rec_hidden=g("rec_hidden"),
        rec_w_dim=g("rec_w_dim"), wv_dim=g("wv_dim"), max_exchange=g("max_exchange"), fixed_exchange=g("fixed_exchange"),
        use_binary=g("use_binary"), entropy_s=g("entropy_s"), entropy_sen=g("entropy_sen"), entropy_rec=g("entropy_rec"),
        first_rec=g("first_rec"), s_prob_prod=g("s_prob_prod"), learning_rate=g("learning_rate"), optim_type=g("optim_type"),
        ignore_receiver=g("ignore_receiver"), batch_global=batch_global, flipout_sen=g("flipout_sen"),
        flipout_rec=g("flipout_rec"), sender_mix=g("sender_mix"), ignore_code=g("ignore_code"), desc_attn=g("desc_attn"),
        desc_attn_dim=g("desc_attn_dim"), n_words=n_words, batch_offset=batch_offset)


def _xavier_normal(rows, cols, gen):
    return torch.zeros(rows, cols).normal_(0.0, math.sqrt(2.0 / (rows + cols)), generator=gen)       # misc.py:367-385


def _default_linear(rows, cols, gen):
    bound = 1.0 / math.sqrt(cols)
    return (torch.zeros(rows, cols).uniform_(-bound, bound, generator=gen), torch.zeros(rows).uniform_(-bound, bound, generator=gen))


def init_params(flags, seed=0):
    """{'receiver', 'sender', 'baseline_rec', 'baseline_sen'} -> OrderedDict(state_dict key -> fp32 tensor)."""
    g = torch.Generator().manual_seed(int(seed))
    F, Hi, M, Hr, WV, Hb = (flags.img_feat_dim, flags.img_h_dim, flags.rec_w_dim, flags.rec_hidden, flags.wv_dim,
                            flags.baseline_hid_dim)
    snd = OrderedDict()
    snd["code_bias"] = torch.zeros(M).normal_(generator=g)                         # model.py:97
    mou = getattr(flags, "sender_mix", "sum") == "mou"
    for name, r, c in (("image_layer", Hi, F), ("code_layer", Hi, M), ("binary_layer", M, 4 * Hi if mou else Hi)):   # model.py:71-76
        snd[name + ".weight"], snd[name + ".bias"] = _xavier_normal(r, c, g), torch.zeros(r)
    if mou and getattr(flags, "ignore_code", False):
        snd["code_bias_mou"] = torch.zeros(M).normal_(generator=g)                 # model.py:73-74
    rec = OrderedDict()
    rec["rnn.weight_ih"], rec["rnn.weight_hh"] = _xavier_normal(3 * Hr, M, g), _xavier_normal(3 * Hr, Hr, g)
    rec["rnn.bias_ih"], rec["rnn.bias_hh"] = torch.zeros(3 * Hr), torch.zeros(3 * Hr)
    rec["w_h.weight"], rec["w_h.bias"] = _xavier_normal(Hr, Hr, g), torch.zeros(Hr)
    rec["w_d.weight"] = _xavier_normal(Hr, WV, g)
    rec["w.weight"], rec["w.bias"] = _xavier_normal(M, Hr, g), torch.zeros(M)
    rec["y1.weight"], rec["y1.bias"] = _xavier_normal(Hr, Hr + WV, g), torch.zeros(Hr)
    rec["y2.weight"], rec["y2.bias"] = _xavier_normal(1, Hr, g), torch.zeros(1)
    rec["s.weight"], rec["s.bias"] = _xavier_normal(1, Hr, g), torch.zeros(1)
    if flags.desc_attn:                                                            # model.py:267-271
        A = flags.desc_attn_dim
        rec["d_d.weight"], rec["d_d.bias"] = _xavier_normal(A, WV, g), torch.zeros(A)
        rec["d_h.weight"], rec["d_h.bias"] = _xavier_normal(A, Hr, g), torch.zeros(A)
        rec["d_attn.weight"], rec["d_attn.bias"] = _xavier_normal(1, A, g), torch.zeros(1)
    out = OrderedDict(receiver=rec, sender=snd)
    for name, width in (("baseline_rec", M + Hr), ("baseline_sen", Hi + M)):
        b = OrderedDict()
        b["linear1.weight"], b["linear1.bias"] = _default_linear(Hb, width, g)
        b["linear2.weight"], b["linear2.bias"] = _default_linear(1, Hb, g)
        out[name] = b
    return out


def batch(flags, seed=0, rows=None):
    """(x (B,F), desc (D,WV), target (B,) int64)."""
    g = torch.Generator().manual_seed(1000 + int(seed))
    B = rows or flags.batch_size
    x = torch.randn(B, flags.img_feat_dim, generator=g)
    desc = torch.randn(flags.n_classes, flags.wv_dim, generator=g)
    target = torch.randint(0, flags.n_classes, (B,), generator=g)
    return x, desc, target


def desc_set(flags, seed=0, min_words=3, max_words=14):
    """-desc_attn: ragged word-level descriptions, dict(desc_set (NW,WV), desc_set_lens [D]); {} when the flag is off.
    The real 30-class set has 3..20 words per class (NW = 259)."""
    if not flags.desc_attn:
        return {}
    g = torch.Generator().manual_seed(2000 + int(seed))
    lens = [int(v) for v in torch.randint(min_words, max_words + 1, (flags.n_classes,), generator=g)]
    return dict(desc_set=torch.randn(sum(lens), flags.wv_dim, generator=g), desc_set_lens=lens)
