"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A PyTorch-CPU fp32 restatement of the reference's referential-game training iteration
(`/root/reference/model.py`), written for Python 3 / torch 2.x with the torch-0.1.12 semantics the
reference was written against made explicit.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module; the product package
`multimodalgame_b200` never does and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors of its own.  This oracle is pinned against
the reference's *own unmodified code* executed in the build container under `oracle/ref_shim.py`
(module forwards, `exchange`, the five loss functions and the update block `model.py:1243-1330`); the
resulting vectors are committed under `tests/golden/` by `tests/golden/make_golden.py` and checked by
`tests/test_oracle_golden.py` (CPU, no reference needed at test time).

Randomness: the reference samples with host `np.random.rand` (float64, `model.py:227,420,460`).  Here
every draw is an *injected* float64 array, consumed in the reference's order per exchange step:
sender message (B,M) -> stop bit (B,1) -> receiver message (B,M).

Structure is deliberately "as written" (image layer recomputed every step, cartesian `build_inp`,
four separate backward passes) so that timing this oracle is a fair stand-in for the reference's CPU
path.
"""
from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-8  # model.py:908-910,919-921


class GameConfig(object):
    """Flag subset that shapes the hot path; names follow the reference's gflags (model.py:1641-1741)."""

    def __init__(self, batch_size=32, img_feat_dim=4096, img_h_dim=100, baseline_hid_dim=500,
                 sender_out_dim=50, rec_hidden=128, rec_w_dim=50, wv_dim=100, n_classes=30,
                 max_exchange=3, fixed_exchange=True, use_binary=True, entropy_s=None,
                 entropy_sen=None, entropy_rec=None, first_rec=0.0, s_prob_prod=True,
                 learning_rate=1e-4, optim_type="RMSprop", top_k_train=6, ignore_receiver=False,
                 flipout_sen=None, flipout_rec=None, sender_mix="sum", ignore_code=False, desc_attn=False,
                 desc_attn_dim=64):
        assert sender_out_dim == rec_w_dim  # model.py:1756
        self.batch_size = batch_size
        self.img_feat_dim = img_feat_dim
        self.img_h_dim = img_h_dim
        self.baseline_hid_dim = baseline_hid_dim
        self.sender_out_dim = sender_out_dim
        self.rec_hidden = rec_hidden
        self.rec_w_dim = rec_w_dim
        self.wv_dim = wv_dim
        self.n_classes = n_classes
        self.max_exchange = max_exchange
        self.fixed_exchange = fixed_exchange
        self.use_binary = use_binary
        self.entropy_s = entropy_s
        self.entropy_sen = entropy_sen
        self.entropy_rec = entropy_rec
        self.first_rec = first_rec
        self.s_prob_prod = s_prob_prod
        self.learning_rate = learning_rate
        self.optim_type = optim_type
        self.top_k_train = top_k_train
        self.ignore_receiver = ignore_receiver
        self.flipout_sen = flipout_sen      # model.py:1710-1711 (None = off)
        self.flipout_rec = flipout_rec
        self.sender_mix = sender_mix        # 'sum' | 'prod' | 'mou' (model.py:1692)
        self.ignore_code = ignore_code      # model.py:1704
        self.desc_attn = desc_attn          # model.py:1719-1720: the receiver attends over the words of every class description
        self.desc_attn_dim = desc_attn_dim

    def as_dict(self):
        return dict(self.__dict__)


# ------------------------------------------------------------------------------------------------
# parameters (state_dict key names and shapes as the reference modules create them)
# ------------------------------------------------------------------------------------------------
def _xavier_normal_(t, gen):
    # misc.py:367-385: N(0, sqrt(2 / (fan_in + fan_out))), fan_in = size(1), fan_out = size(0)
    std = math.sqrt(2.0 / (t.shape[0] + t.shape[1]))
    return t.normal_(0.0, std, generator=gen)


def _torch_default_linear_(w, b, gen):
    # Baseline keeps nn.Linear's default init (model.py:480-494 has no reset_parameters):
    # U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias.
    bound = 1.0 / math.sqrt(w.shape[1])
    w.uniform_(-bound, bound, generator=gen)
    b.uniform_(-bound, bound, generator=gen)


def init_params(cfg, seed=0):
    """Returns {'receiver','sender','baseline_rec','baseline_sen'} -> OrderedDict(name -> fp32 tensor).

    Key names/shapes: model.py:67-76 (Sender), 256-265 (Receiver), 492-494 (Baseline); init:
    model.py:90-97, 275-288, misc.py:367-385.  The random stream is this oracle's own (torch Generator);
    parity tests always pass explicit parameter values to both implementations."""
    g = torch.Generator().manual_seed(seed)
    Fd, Hi, M, Hr, WV, Hb = (cfg.img_feat_dim, cfg.img_h_dim, cfg.rec_w_dim, cfg.rec_hidden, cfg.wv_dim,
                             cfg.baseline_hid_dim)
    z = lambda *s: torch.zeros(*s, dtype=torch.float32)
    snd = OrderedDict()
    snd["code_bias"] = z(M).normal_(generator=g)
    snd["image_layer.weight"] = _xavier_normal_(z(Hi, Fd), g)
    snd["image_layer.bias"] = z(Hi)
    snd["code_layer.weight"] = _xavier_normal_(z(Hi, M), g)
    snd["code_layer.bias"] = z(Hi)
    mou = cfg.sender_mix == "mou"
    snd["binary_layer.weight"] = _xavier_normal_(z(M, 4 * Hi if mou else Hi), g)   # model.py:71-76
    snd["binary_layer.bias"] = z(M)
    if mou and cfg.ignore_code:
        # model.py:73-74 creates the parameter and reset_parameters (model.py:90-97) leaves it uninitialised (torch.Tensor(n));
        # N(0, 1) like code_bias is this package's choice, any finite value is equally "reference"
        snd["code_bias_mou"] = z(M).normal_(generator=g)
    rec = OrderedDict()
    rec["rnn.weight_ih"] = _xavier_normal_(z(3 * Hr, M), g)
    rec["rnn.weight_hh"] = _xavier_normal_(z(3 * Hr, Hr), g)
    rec["rnn.bias_ih"] = z(3 * Hr)
    rec["rnn.bias_hh"] = z(3 * Hr)
    rec["w_h.weight"] = _xavier_normal_(z(Hr, Hr), g)
    rec["w_h.bias"] = z(Hr)
    rec["w_d.weight"] = _xavier_normal_(z(Hr, WV), g)
    rec["w.weight"] = _xavier_normal_(z(M, Hr), g)
    rec["w.bias"] = z(M)
    rec["y1.weight"] = _xavier_normal_(z(Hr, Hr + WV), g)
    rec["y1.bias"] = z(Hr)
    rec["y2.weight"] = _xavier_normal_(z(1, Hr), g)
    rec["y2.bias"] = z(1)
    rec["s.weight"] = _xavier_normal_(z(1, Hr), g)
    rec["s.bias"] = z(1)
    if cfg.desc_attn:                                                             # model.py:267-271
        A = cfg.desc_attn_dim
        rec["d_d.weight"] = _xavier_normal_(z(A, WV), g)
        rec["d_d.bias"] = z(A)
        rec["d_h.weight"] = _xavier_normal_(z(A, Hr), g)
        rec["d_h.bias"] = z(A)
        rec["d_attn.weight"] = _xavier_normal_(z(1, A), g)
        rec["d_attn.bias"] = z(1)
    bsen = OrderedDict()
    bsen["linear1.weight"], bsen["linear1.bias"] = z(Hb, Hi + M), z(Hb)
    _torch_default_linear_(bsen["linear1.weight"], bsen["linear1.bias"], g)
    bsen["linear2.weight"], bsen["linear2.bias"] = z(1, Hb), z(1)
    _torch_default_linear_(bsen["linear2.weight"], bsen["linear2.bias"], g)
    brec = OrderedDict()
    brec["linear1.weight"], brec["linear1.bias"] = z(Hb, M + Hr), z(Hb)
    _torch_default_linear_(brec["linear1.weight"], brec["linear1.bias"], g)
    brec["linear2.weight"], brec["linear2.bias"] = z(1, Hb), z(1)
    _torch_default_linear_(brec["linear2.weight"], brec["linear2.bias"], g)
    return OrderedDict(receiver=rec, sender=snd, baseline_rec=brec, baseline_sen=bsen)


AGENTS = ("receiver", "sender", "baseline_rec", "baseline_sen")  # optimizer order, model.py:1308-1330


def clone_params(params, requires_grad=False):
    out = OrderedDict()
    for a in params:
        out[a] = OrderedDict((k, v.detach().clone().requires_grad_(requires_grad)) for k, v in params[a].items())
    return out


# ------------------------------------------------------------------------------------------------
# sampling
# ------------------------------------------------------------------------------------------------
def _bernoulli(probs, u):
    """`(np.random.rand(*shape) < probs_).astype('float32')` — float64 uniform vs float32 prob promoted
    to float64 (model.py:225-227)."""
    p = probs.detach().numpy().astype(np.float64)
    return torch.from_numpy((np.asarray(u, dtype=np.float64) < p).astype(np.float32))


# ------------------------------------------------------------------------------------------------
# agents
# ------------------------------------------------------------------------------------------------
def flipout(binary, p, u):
    """model.py:554-568: |bit - 1[u < p]| with its own uniform draw."""
    mask = torch.from_numpy((np.asarray(u, dtype=np.float64) < float(p)).astype("float32"))
    return (binary - mask).abs()


def sender_forward(P, x, w, t, cfg, train, u=None, u_flip=None):
    """Sender.forward default path (sender_mix='sum', no attention): model.py:195-238.
    Returns (message, probs_or_None, h_x)."""
    h_x = F.linear(x, P["image_layer.weight"], P["image_layer.bias"])          # :195
    if t == 0:
        first_code = torch.sigmoid(P["code_bias"].view(1, -1))                  # :199
        h_w = F.linear(first_code, P["code_layer.weight"], P["code_layer.bias"]).expand(x.shape[0], -1)
    elif cfg.ignore_code and cfg.sender_mix == "mou":
        code_mou = torch.sigmoid(P["code_bias_mou"].view(1, -1))               # :201-205: the same learned code for every example
        h_w = F.linear(code_mou, P["code_layer.weight"], P["code_layer.bias"]).expand(x.shape[0], -1)
    else:
        h_w = F.linear(w, P["code_layer.weight"], P["code_layer.bias"])        # :207
    assert cfg.sender_mix in ("sum", "prod", "mou")
    if cfg.sender_mix == "mou":
        mixed = torch.cat([h_x, h_w, h_x - h_w, h_x * h_w], 1)                   # :211-213, 219-221 (with and without ignore_code)
    elif cfg.ignore_code:
        mixed = h_x                                                             # :208-210
    elif cfg.sender_mix == "prod":
        mixed = h_x * h_w                                                       # :217-218
    else:
        mixed = h_x + h_w                                                       # :215-216
    feats = F.linear(torch.tanh(mixed), P["binary_layer.weight"], P["binary_layer.bias"])
    if cfg.use_binary:
        probs = torch.sigmoid(feats)                                            # :223
        if train:
            msg = _bernoulli(probs, u)                                          # :225-227
        else:
            msg = torch.round(probs).detach()                                   # :229
        if cfg.flipout_sen is not None and train:                               # :233-234 (flipout_dev not modelled)
            msg = flipout(msg, cfg.flipout_sen, u_flip)
        return msg, probs, h_x
    return feats, None, h_x                                                     # :238


def build_inp(h, desc):
    """Cartesian product rows [h_b ; desc_d], b-major (model.py:519-551)."""
    B, D = h.shape[0], desc.shape[0]
    hb = h.unsqueeze(1).expand(B, D, h.shape[1]).reshape(B * D, -1)
    dd = desc.unsqueeze(0).expand(B, D, desc.shape[1]).reshape(B * D, -1)
    return torch.cat([hb, dd], 1)


def gru_cell(P, z, h):
    """nn.GRUCell, gate order r,z,n (model.py:256,340)."""
    gi = F.linear(z, P["rnn.weight_ih"], P["rnn.bias_ih"])
    gh = F.linear(h, P["rnn.weight_hh"], P["rnn.bias_hh"])
    i_r, i_u, i_n = gi.chunk(3, 1)
    h_r, h_u, h_n = gh.chunk(3, 1)
    r = torch.sigmoid(i_r + h_r)
    u = torch.sigmoid(i_u + h_u)
    n = torch.tanh(i_n + r * h_n)
    return n + u * (h - n)


def attend_descriptions(P, h_z, desc_set, desc_set_lens):
    """-desc_attn (model.py:344-410): additive attention of the hidden state over the NW description words,
    softmax within each class's word segment, per-(example, class) weighted bag of words.
    Returns (B, D, WV)."""
    dd = F.linear(desc_set, P["d_d.weight"], P["d_d.bias"])                       # :352  NW x A
    dh = F.linear(h_z, P["d_h.weight"], P["d_h.bias"])                            # :359  B x A
    e = F.linear(torch.tanh(dd.unsqueeze(0) + dh.unsqueeze(1)), P["d_attn.weight"], P["d_attn.bias"]).squeeze(2)  # :366
    out, start = [], 0
    for n in desc_set_lens:                                                       # :372-397
        a = torch.softmax(e[:, start:start + n], 1)                               # B x NW_i
        out.append((a.unsqueeze(2) * desc_set[start:start + n].unsqueeze(0)).sum(1, keepdim=True))
        start += n
    return torch.cat(out, 1)


def receiver_forward(P, z, desc, state, cfg, train, u_s=None, u_w=None, u_flip=None, desc_set=None,
                     desc_set_lens=None):
    """Receiver.forward: model.py:333-342, 412-477; -desc_attn branch 344-410.
    `state` = dict(h_z, s_prob_prod) mutated like the module attributes.
    Returns ((s_binary, s_prob), (w_feats, w_probs), y)."""
    B = z.shape[0]
    if state.get("h_z") is None:
        state["h_z"] = torch.zeros(B, cfg.rec_hidden)                            # :336-337
    h_z = state["h_z"] = gru_cell(P, z, state["h_z"])                            # :340
    if cfg.desc_attn:
        weighted_desc = attend_descriptions(P, h_z, desc_set, desc_set_lens)     # B x D x WV
        D = weighted_desc.shape[1]
        inp = torch.cat([weighted_desc.reshape(B * D, -1),
                         h_z.unsqueeze(1).expand(B, D, h_z.shape[1]).reshape(B * D, -1)], 1)   # :408-410 [desc ; h_z]
    else:
        inp = build_inp(h_z, desc)                                               # :412
    s_prob = torch.sigmoid(F.linear(h_z, P["s.weight"], P["s.bias"]))            # :414-415
    if train:
        s_binary = _bernoulli(s_prob, u_s)                                       # :418-420
    else:
        if state.get("s_prob_prod") is None or not cfg.s_prob_prod:              # :423-426
            state["s_prob_prod"] = s_prob
        else:
            state["s_prob_prod"] = state["s_prob_prod"] * s_prob
        s_binary = torch.round(state["s_prob_prod"]).detach()                    # :427
    y = F.linear(inp, P["y1.weight"], P["y1.bias"]).clamp(min=0)                 # :432
    y = F.linear(y, P["y2.weight"], P["y2.bias"]).view(B, -1)                    # :433
    y_scores = torch.softmax(y, 1).detach()                                      # :441
    wd_src = weighted_desc if cfg.desc_attn else desc.unsqueeze(0)               # :444-448 (not detached)
    wd_inp = (y_scores.unsqueeze(2) * wd_src).sum(1)                             # :442-449
    h_w = torch.tanh(F.linear(h_z, P["w_h.weight"], P["w_h.bias"]) + F.linear(wd_inp, P["w_d.weight"]))  # :452
    state["h_w"] = h_w
    w_scores = F.linear(h_w, P["w.weight"], P["w.bias"])                         # :454
    if cfg.use_binary:
        w_probs = torch.sigmoid(w_scores)
        if train:
            w_feats = _bernoulli(w_probs, u_w)                                   # :458-460
        else:
            w_feats = torch.round(w_probs).detach()                              # :462
        if cfg.flipout_rec is not None and train:                                # :467-468
            w_feats = flipout(w_feats, cfg.flipout_rec, u_flip)
        if cfg.ignore_receiver:
            w_feats = torch.zeros_like(w_feats)                                  # :470-472
    else:
        w_feats, w_probs = w_scores, None                                        # :474-475
    return (s_binary, s_prob), (w_feats, w_probs), y


def baseline_forward(P, x, binary, inp):
    """Baseline.forward: model.py:496-516."""
    feats = torch.cat([t for t in (x, binary, inp) if t is not None], 1)
    hidden = F.linear(feats, P["linear1.weight"], P["linear1.bias"]).clamp(min=0)
    return F.linear(hidden, P["linear2.weight"], P["linear2.bias"])


# ------------------------------------------------------------------------------------------------
# exchange (model.py:725-876)
# ------------------------------------------------------------------------------------------------
def exchange(params, x, desc, cfg, train, uniforms=None, break_early=False, corrupt_mask=None, desc_set=None,
             desc_set_lens=None):
    """`uniforms`: per step a triple (u_z (B,M), u_s (B,1), u_w (B,M)) of float64 arrays (train only).
    `corrupt_mask`: optional (M,) 0/1 tensor XOR-ed onto the sender message (model.py:814-820).
    Returns a dict with the reference's lists: stop_mask[T'+1] (uint8), stop_feat, stop_prob, sen_feats,
    sen_probs, rec_feats, rec_probs, y, bs, br, plus h_x / per-step h_z for kernel-level checks."""
    B = x.shape[0]
    S, R = params["sender"], params["receiver"]
    out = dict(stop_mask=[torch.ones(B, 1, dtype=torch.uint8)], stop_feat=[], stop_prob=[], sen_feats=[],
               sen_probs=[], rec_feats=[], rec_probs=[], y=[], bs=[], br=[], h_z=[], h_w=[])
    w_binary = torch.full((B, cfg.rec_w_dim), float(cfg.first_rec))                # :786
    state = dict(h_z=None, s_prob_prod=None)                                        # :798-799
    for t in range(cfg.max_exchange):                                               # :801
        z_r = w_binary
        u = tuple(uniforms[t]) if train else (None, None, None)
        u = u + (None,) * (5 - len(u))          # (u_z, u_s, u_w [, u_flip_z, u_flip_w])
        z_binary, z_probs, h_x = sender_forward(S, x, z_r.detach(), t, cfg, train, u[0], u[3])   # :807-811
        if corrupt_mask is not None:
            z_binary = (z_binary - corrupt_mask.view(1, -1)).abs()                  # :814-820
        (s_binary, s_prob), (w_binary, w_probs), outp = receiver_forward(
            R, z_binary.detach(), desc.detach(), state, cfg, train, u[1], u[2], u[4], desc_set, desc_set_lens)  # :826-829
        if train:
            out["bs"].append(baseline_forward(params["baseline_sen"], h_x.detach(), z_r.detach(), None))   # :835-836
            out["br"].append(baseline_forward(params["baseline_rec"], None, z_binary.detach(),
                                              state["h_z"].detach()))                                        # :842-843
        out["stop_mask"].append(torch.min(out["stop_mask"][-1], s_binary.to(torch.uint8)))  # :852
        out["stop_feat"].append(s_binary)
        out["stop_prob"].append(s_prob)
        out["sen_feats"].append(z_binary)
        out["sen_probs"].append(z_probs)
        out["rec_feats"].append(w_binary)
        out["rec_probs"].append(w_probs)
        out["y"].append(outp)
        out["h_z"].append(state["h_z"])
        out["h_w"].append(state["h_w"])
        out["h_x"] = h_x
        if break_early and float(out["stop_mask"][-1].float().sum()) == 0:          # :866
            break
    out["stop_mask"][-1] = torch.zeros_like(out["stop_mask"][-1])                   # :870
    return out


# ------------------------------------------------------------------------------------------------
# losses (model.py:879-988)
# ------------------------------------------------------------------------------------------------
def get_rec_outp(y, masks):
    """model.py:879-904.  masks: list of (B,1) uint8 one-hot-over-steps, or None (fixed)."""
    negent = [(torch.log(torch.softmax(yy, 1) + EPS) * torch.softmax(yy, 1)).sum(1).mean() for yy in y]
    if masks is None:
        return y[-1], negent
    B = y[0].shape[0]
    inp = torch.stack(y, 1)                                 # (B, T, D)
    m = torch.cat(masks, 1).bool()                          # (B, T)
    assert bool((m.sum(1) == 1).all()), "each example must stop exactly once (model.py:898-900)"
    return inp[m].view(B, -1), negent


def calculate_loss_binary(feats, probs, logs, baseline_scores, entropy_penalty):
    """model.py:907-927 with 0.1.12 shapes: every per-example quantity is (B,1); no broadcasting."""
    f = feats.detach()
    log_p_z = (f * torch.log(probs + EPS) + (1 - f) * torch.log(1 - probs + EPS)).sum(1, keepdim=True)
    weight = logs.detach() - baseline_scores.detach()
    if logs.shape[0] > 1:
        weight = weight / max(1.0, float(torch.std(weight)))     # unbiased std over all elements
    loss = torch.mean(-1 * weight * log_p_z)
    initial_negent = (torch.log(probs + EPS) * probs).sum(1).mean()
    inverse_negent = (torch.log((1.0 - probs) + EPS) * (1.0 - probs)).sum(1).mean()
    negentropy = initial_negent + inverse_negent
    if entropy_penalty is not None:
        loss = loss + entropy_penalty * negentropy
    return loss, negentropy


def multistep_loss_binary(feats, probs, logs, baseline_scores, masks, entropy_penalty):
    """model.py:930-968."""
    if masks is not None:
        sums = [float(m.float().sum()) for m in masks]
        losses, ents = [], []
        for f, p, b, m, ms in zip(feats, probs, baseline_scores, masks, sums):
            if ms == 0:
                losses.append(torch.zeros(()))
                continue
            sel = m.view(-1).bool()
            l, e = calculate_loss_binary(f[sel], p[sel], logs[sel], b[sel], entropy_penalty)
            losses.append(l)
            ents.append(e)
        loss = sum(l * ms for l, ms in zip(losses, sums)) / sum(sums)
    else:
        pairs = [calculate_loss_binary(f, p, logs, b, entropy_penalty)
                 for f, p, b in zip(feats, probs, baseline_scores)]
        losses = [o[0] for o in pairs]
        ents = [o[1] for o in pairs]
        loss = sum(losses) / len(feats)
    return loss, ents


def calculate_loss_bas(baseline_scores, logs):
    return F.mse_loss(baseline_scores, logs.detach())             # model.py:971-973


def multistep_loss_bas(baseline_scores, logs, masks):
    """model.py:976-988."""
    if masks is not None:
        losses, sums = [], []
        for b, m in zip(baseline_scores, masks):
            sel = m.view(-1).bool()
            ms = float(m.float().sum())
            sums.append(ms)
            losses.append(calculate_loss_bas(b[sel].view(-1, 1), logs[sel].view(-1, 1)) if ms > 0
                          else torch.zeros(()))
        return sum(l * ms for l, ms in zip(losses, sums)) / sum(sums)
    losses = [calculate_loss_bas(b, logs) for b in baseline_scores]
    return sum(losses) / len(baseline_scores)


def compute_losses(ex, target, cfg):
    """Mask wiring and loss assembly of run(): model.py:1243-1305.  Returns a dict of tensors."""
    s_masks = ex["stop_mask"]
    if cfg.fixed_exchange:
        m_s = m_rec = m_sen = m_brec = m_bsen = y_masks = None
    else:
        m_s = s_masks[:-1]
        m_rec = s_masks[1:-1]
        m_sen = s_masks[:-1]
        m_brec = s_masks[:-1]
        m_bsen = s_masks[:-1]
        y_masks = [torch.min(1 - m1, m2) for m1, m2 in zip(s_masks[1:], s_masks[:-1])]
    outp, ent_y = get_rec_outp(ex["y"], y_masks)
    dist = torch.log_softmax(outp, 1)
    argmax = dist.detach().argmax(1)
    nll = F.nll_loss(dist, target)
    logs = dist.detach().gather(1, target.view(-1, 1))
    res = dict(outp=outp, dist=dist, argmax=argmax, nll_loss=nll, logs=logs, ent_y=ent_y)
    zero = torch.zeros(())
    if cfg.use_binary:
        if not cfg.fixed_exchange:
            res["loss_binary_s"], res["ent_binary_s"] = multistep_loss_binary(
                ex["stop_feat"], ex["stop_prob"], logs, ex["br"], m_s, cfg.entropy_s)
        if len(ex["rec_feats"][:-1]) > 0:
            res["loss_binary_rec"], res["ent_binary_rec"] = multistep_loss_binary(
                ex["rec_feats"][:-1], ex["rec_probs"][:-1], logs, ex["br"][:-1], m_rec, cfg.entropy_rec)
        else:
            res["loss_binary_rec"], res["ent_binary_rec"] = zero, []
        res["loss_binary_sen"], res["ent_binary_sen"] = multistep_loss_binary(
            ex["sen_feats"], ex["sen_probs"], logs, ex["bs"], m_sen, cfg.entropy_sen)
        res["loss_bas_rec"] = multistep_loss_bas(ex["br"], logs, m_brec)
        res["loss_bas_sen"] = multistep_loss_bas(ex["bs"], logs, m_bsen)
    loss_rec = nll
    if cfg.use_binary:
        loss_rec = loss_rec + res["loss_binary_rec"]
        if not cfg.fixed_exchange:
            loss_rec = loss_rec + res["loss_binary_s"]
        res["loss_sen"] = res["loss_binary_sen"]
    else:
        res["loss_sen"] = res["loss_bas_rec"] = res["loss_bas_sen"] = zero
    res["loss_rec"] = loss_rec
    return res


def topk_accuracy(dist, target, k):
    """model.py:1333-1338: target within the last k of the ascending argsort."""
    top = np.argsort(dist.detach().numpy(), axis=1)[:, -k:]
    return float((top == target.numpy().reshape(-1, 1)).sum()) / float(target.shape[0])


# ------------------------------------------------------------------------------------------------
# optimizer step (model.py:1308-1330; torch.nn.utils.clip_grad_norm, torch.optim.{RMSprop,Adam,SGD})
# ------------------------------------------------------------------------------------------------
def clip_grad_norm(grads, max_norm=1.0):
    """total = sqrt(sum ||g||^2); coef = max_norm / (total + 1e-6); scale if coef < 1.  Returns total."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    coef = max_norm / (total + 1e-6)
    if coef < 1:
        for g in grads:
            g.mul_(coef)
    return total


def new_opt_state(params):
    st = OrderedDict()
    for a in params:
        st[a] = OrderedDict((k, dict(step=0, sq=torch.zeros_like(v), m=torch.zeros_like(v)))
                            for k, v in params[a].items())
    return st


def optimizer_step(p, g, st, cfg):
    lr = cfg.learning_rate
    if cfg.optim_type == "RMSprop":       # alpha=0.99, eps=1e-8, no momentum, not centered
        st["sq"].mul_(0.99).addcmul_(g, g, value=0.01)
        p.addcdiv_(g, st["sq"].sqrt().add_(1e-8), value=-lr)
    elif cfg.optim_type == "SGD":
        p.add_(g, alpha=-lr)
    elif cfg.optim_type == "Adam":        # betas (0.9, 0.999), eps 1e-8; torch>=1.x form (denominator /sqrt(bc2) + eps)
        st["step"] += 1
        st["m"].mul_(0.9).add_(g, alpha=0.1)
        st["sq"].mul_(0.999).addcmul_(g, g, value=0.001)
        bc1 = 1 - 0.9 ** st["step"]
        bc2 = 1 - 0.999 ** st["step"]
        denom = (st["sq"].sqrt() / math.sqrt(bc2)).add_(1e-8)
        p.addcdiv_(st["m"], denom, value=-lr / bc1)
    else:
        raise NotImplementedError(cfg.optim_type)


def train_iteration(params, opt_state, x, target, desc, cfg, uniforms, return_grads=False, desc_set=None,
                    desc_set_lens=None):
    """One iteration of run()'s loop body: model.py:1229-1339.  `params` (leaf tensors) are updated in
    place.  Returns (exchange dict, losses dict[, grads])."""
    for a in params:
        for v in params[a].values():
            v.requires_grad_(True)
            v.grad = None
    ex = exchange(params, x, desc, cfg, True, uniforms, break_early=not cfg.fixed_exchange, desc_set=desc_set,
                  desc_set_lens=desc_set_lens)
    res = compute_losses(ex, target, cfg)
    plan = [("receiver", "loss_rec")]
    if cfg.use_binary:
        plan += [("sender", "loss_sen"), ("baseline_rec", "loss_bas_rec"), ("baseline_sen", "loss_bas_sen")]
    grads = OrderedDict()
    norms = OrderedDict()
    for agent, lname in plan:
        names = list(params[agent].keys())
        gs = torch.autograd.grad(res[lname], [params[agent][k] for k in names], allow_unused=True,
                                 retain_graph=True)
        live = [(k, g.clone()) for k, g in zip(names, gs) if g is not None]
        grads[agent] = OrderedDict((k, (g.clone() if g is not None else None)) for k, g in zip(names, gs))
        norms[agent] = clip_grad_norm([g for _, g in live], 1.0)
        with torch.no_grad():
            for k, g in live:
                optimizer_step(params[agent][k], g, opt_state[agent][k], cfg)
    res["grad_norms"] = norms
    res["accuracy"] = topk_accuracy(res["dist"], target, min(cfg.top_k_train, cfg.n_classes))
    for a in params:
        for v in params[a].values():
            v.requires_grad_(False)
    if return_grads:
        return ex, res, grads
    return ex, res


def draw_uniforms(rng, cfg, B=None, steps=None):
    """Uniforms in the reference's consumption order from a numpy RandomState."""
    B = B or cfg.batch_size
    M = cfg.rec_w_dim
    out = []
    for _ in range(steps or cfg.max_exchange):      # reference draw order: z, [flip z], s, w, [flip w]
        u_z = rng.rand(B, M)
        u_fz = rng.rand(B, M) if cfg.flipout_sen is not None else None
        u_s = rng.rand(B, 1)
        u_w = rng.rand(B, M)
        u_fw = rng.rand(B, M) if cfg.flipout_rec is not None else None
        out.append((u_z, u_s, u_w, u_fz, u_fw) if (u_fz is not None or u_fw is not None) else (u_z, u_s, u_w))
    return out


def synthetic_desc_set(cfg, seed=0, min_words=1, max_words=12):
    """Ragged word-level descriptions for -desc_attn: (desc_set (NW, WV), desc_set_lens [D]); the real data has
    3..20 words per class (NW = 259 for the 30-class set)."""
    g = torch.Generator().manual_seed(2000 + seed)
    lens = [int(v) for v in torch.randint(min_words, max_words + 1, (cfg.n_classes,), generator=g)]
    return torch.randn(sum(lens), cfg.wv_dim, generator=g), lens


def synthetic_batch(cfg, seed=0, B=None):
    """x ~ N(0,1) (B,F); desc ~ N(0,1) (D,WV) as wv_type=fake (model.py:1069); target ~ U{0..D-1}."""
    g = torch.Generator().manual_seed(1000 + seed)
    B = B or cfg.batch_size
    x = torch.randn(B, cfg.img_feat_dim, generator=g)
    desc = torch.randn(cfg.n_classes, cfg.wv_dim, generator=g)
    target = torch.randint(0, cfg.n_classes, (B,), generator=g)
    return x, desc, target
