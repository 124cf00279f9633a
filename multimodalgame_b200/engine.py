"""GameEngine — owns the device buffers of one referential game (flat parameters / gradients / optimizer state /
workspace) and drives the C-ABI entry points (include/mmg_b200.h).  PyTorch is used for device memory, streams and
`torch.distributed` only; every tensor operation of the path runs in libmmg_b200.so.

The reference keeps four `nn.Module`s and four `torch.optim` objects (model.py:1014-1142); here their parameters are
views into one flat fp32 buffer so that clip + optimizer step is one kernel and a data-parallel run all-reduces one
contiguous gradient buffer.
"""
import ctypes as C
import os
from collections import OrderedDict

import numpy as np
import torch

from . import capi


def make_config(batch, n_classes, img_feat_dim=4096, img_h_dim=100, baseline_hid_dim=500, sender_out_dim=50,
                rec_hidden=128, rec_w_dim=50, wv_dim=100, max_exchange=3, fixed_exchange=True, use_binary=True,
                entropy_s=None, entropy_sen=None, entropy_rec=None, first_rec=0.0, s_prob_prod=True,
                learning_rate=1e-4, optim_type="RMSprop", ignore_receiver=False, batch_global=None, max_norm=1.0,
                flipout_sen=None, flipout_rec=None, flipout_dev=False, sender_mix="sum", ignore_code=False,
                desc_attn=False, desc_attn_dim=64, n_words=0, batch_offset=0):
    """Build the C config from reference flag names/defaults (model.py:1641-1741)."""
    assert sender_out_dim == rec_w_dim, \
        "Both sender and receiver should communicate with same dim vectors for now."   # model.py:1756
    c = capi.Config()
    c.batch = int(batch)
    c.batch_global = int(batch_global if batch_global is not None else batch)
    c.img_feat_dim, c.img_h_dim, c.msg_dim = int(img_feat_dim), int(img_h_dim), int(rec_w_dim)
    c.rec_hidden, c.n_classes, c.wv_dim = int(rec_hidden), int(n_classes), int(wv_dim)
    c.baseline_hid, c.max_exchange = int(baseline_hid_dim), int(max_exchange)
    c.use_binary, c.fixed_exchange, c.s_prob_prod = int(bool(use_binary)), int(bool(fixed_exchange)), int(bool(s_prob_prod))
    if optim_type not in capi.OPTIM:
        raise NotImplementedError(optim_type)                                         # model.py:1137
    c.optim_type = capi.OPTIM[optim_type]
    c.has_entropy_s, c.entropy_s = int(entropy_s is not None), float(entropy_s or 0.0)
    c.has_entropy_sen, c.entropy_sen = int(entropy_sen is not None), float(entropy_sen or 0.0)
    c.has_entropy_rec, c.entropy_rec = int(entropy_rec is not None), float(entropy_rec or 0.0)
    c.first_rec, c.learning_rate, c.max_norm = float(first_rec), float(learning_rate), float(max_norm)
    c.ignore_receiver = int(bool(ignore_receiver))
    c.has_flipout_sen, c.flipout_sen = int(flipout_sen is not None), float(flipout_sen or 0.0)     # model.py:1710-1712
    c.has_flipout_rec, c.flipout_rec = int(flipout_rec is not None), float(flipout_rec or 0.0)
    c.flipout_dev = int(bool(flipout_dev))
    if sender_mix not in capi.SENDER_MIX:
        raise NotImplementedError("sender_mix=%s (model.py:1692 knows sum, prod, mou)" % sender_mix)
    c.sender_mix, c.ignore_code = capi.SENDER_MIX[sender_mix], int(bool(ignore_code))
    c.desc_attn = int(bool(desc_attn))                                                               # model.py:1719-1720
    c.desc_attn_dim, c.n_words = (int(desc_attn_dim), int(n_words)) if desc_attn else (0, 0)
    c.batch_offset = int(batch_offset)      # rank * batch in a data-parallel run: the sampler is keyed by the global row
    return c


def _is_vector(key):
    return key.endswith("bias") or key.endswith("bias_ih") or key.endswith("bias_hh") or key == "code_bias_mou"


class GameEngine(object):
    def __init__(self, config, device=None, lib=None, seed=0):
        self.lib = lib if lib is not None else capi.load()
        self.cfg = config
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.layout = capi.ParamLayout()
        self.lib.call("mmg_param_layout_get", C.byref(self.cfg), C.byref(self.layout))
        self.wsl = capi.WorkspaceLayout()
        self.lib.call("mmg_workspace_layout_get", C.byref(self.cfg), C.byref(self.wsl))
        n = int(self.layout.total)
        self.params = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.state1 = torch.zeros(n, dtype=torch.float32, device=self.device)     # RMSprop square_avg / Adam exp_avg_sq
        self.state2 = torch.zeros(n, dtype=torch.float32, device=self.device) if self.cfg.optim_type == 1 else None
        self.workspace = torch.zeros(int(self.wsl.total_bytes), dtype=torch.uint8, device=self.device)
        self.step = 0
        self._keep = None
        self.lib.call("mmg_workspace_init", C.byref(self.cfg), self.workspace.data_ptr(), C.c_uint64(seed), self._stream())
        self._views = {}

    # ---- plumbing ---------------------------------------------------------------------------------------------
    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    @property
    def dims(self):
        c = self.cfg
        return dict(B=c.batch, F=c.img_feat_dim, Hi=c.img_h_dim, M=c.msg_dim, Hr=c.rec_hidden, D=c.n_classes,
                    WV=c.wv_dim, Hb=c.baseline_hid, T=c.max_exchange)

    def _flat_view(self, flat, idx):
        off, rows, cols = int(self.layout.offset[idx]), int(self.layout.rows[idx]), int(self.layout.cols[idx])
        v = flat[off:off + rows * cols]
        agent, key = capi.PARAM_NAMES[idx]
        return v if _is_vector(key) else v.view(rows, cols)

    def named_views(self, flat=None):
        """{agent: OrderedDict(state_dict key -> view into the flat buffer)} (reference key names)."""
        flat = self.params if flat is None else flat
        out = OrderedDict((a, OrderedDict()) for a in capi.SEGMENTS)
        for idx, (agent, key) in enumerate(capi.PARAM_NAMES):
            if int(self.layout.rows[idx]) * int(self.layout.cols[idx]) == 0:
                continue                                   # tensors of a branch that is switched off (-desc_attn)
            out[agent][key] = self._flat_view(flat, idx)
        return out

    def param_keys(self, agent):
        """state_dict keys of one module, in the flat buffer's (= the reference's registration) order."""
        return list(self.named_views()[agent].keys())

    def set_desc_set(self, desc_set, desc_set_lens):
        """-desc_attn: the words of all class descriptions (NW, WV) and the word count of each class (model.py:765-766)."""
        dev = self.device
        lens = torch.as_tensor([int(v) for v in desc_set_lens], dtype=torch.int32)
        assert lens.numel() == self.cfg.n_classes and int(lens.min()) >= 1, "one non-empty word segment per class"
        assert int(lens.sum()) == self.cfg.n_words, (int(lens.sum()), self.cfg.n_words)
        ds = torch.as_tensor(desc_set).to(device=dev, dtype=torch.float32).contiguous()
        assert tuple(ds.shape) == (self.cfg.n_words, self.cfg.wv_dim), tuple(ds.shape)
        self._desc_words = (ds, lens.to(dev))

    def load_params(self, params):
        """params: {agent: {key: tensor}} (e.g. four state_dicts)."""
        views = self.named_views()
        with torch.no_grad():
            for agent in params:
                for key, val in params[agent].items():
                    views[agent][key].copy_(torch.as_tensor(val).to(self.device).reshape(views[agent][key].shape))

    def ws(self, name, shape, dtype=torch.float32):
        """Typed view of a named workspace array."""
        k = (name, tuple(shape), dtype)
        v = self._views.get(k)
        if v is None:
            off = int(getattr(self.wsl, name))
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            v = self.workspace[off:off + nbytes].view(dtype).view(*shape)
            self._views[k] = v
        return v

    def _inputs(self, x, desc, target, train, uniforms=None, corrupt_mask=None, h0=None, top_k=6):
        dev = self.device
        def prep(t, dtype):
            if t is None:
                return None
            t = torch.as_tensor(t)
            if t.dtype != dtype or t.device != dev or not t.is_contiguous():
                t = t.to(device=dev, dtype=dtype).contiguous()
            return t
        keep = [prep(x, torch.float32), prep(desc, torch.float32), prep(target, torch.int64)]
        us = [None, None, None, None, None]       # sender, stop, receiver [, sender flipout, receiver flipout]
        if uniforms is not None:
            us = [prep(u, torch.float64) for u in uniforms] + [None] * (5 - len(uniforms))
        keep += us[:3] + [prep(corrupt_mask, torch.float32), prep(h0, torch.float32)] + us[3:]
        d = self.dims
        assert tuple(keep[0].shape) == (d["B"], d["F"]), (tuple(keep[0].shape), (d["B"], d["F"]))
        assert tuple(keep[1].shape) == (d["D"], d["WV"]), tuple(keep[1].shape)
        inp = capi.Inputs()
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        inp.d_x, inp.d_desc, inp.d_target = ptr(keep[0]), ptr(keep[1]), ptr(keep[2])
        inp.d_u_sen, inp.d_u_stop, inp.d_u_rec = ptr(keep[3]), ptr(keep[4]), ptr(keep[5])
        inp.d_corrupt_mask, inp.d_h0 = ptr(keep[6]), ptr(keep[7])
        inp.d_u_flip_sen, inp.d_u_flip_rec = ptr(keep[8]), ptr(keep[9])
        inp.top_k, inp.train = int(top_k), int(bool(train))
        if self.cfg.desc_attn:
            words = getattr(self, "_desc_words", None)
            if words is None:
                raise ValueError("-desc_attn: call set_desc_set(desc_set, desc_set_lens) first")
            inp.d_desc_set, inp.d_desc_set_lens = ptr(words[0]), ptr(words[1])
            keep = keep + list(words)
        self._keep = keep      # keep the tensors alive until the next call
        return inp

    # ---- the path ------------------------------------------------------------------------------------------------
    def forward(self, x, desc, target=None, train=True, uniforms=None, corrupt_mask=None, h0=None, top_k=6):
        self._inp = self._inputs(x, desc, target, train, uniforms, corrupt_mask, h0, top_k)
        self.lib.call("mmg_exchange_forward", C.byref(self.cfg), self.params.data_ptr(), C.byref(self._inp),
                      self.workspace.data_ptr(), self._stream())

    def loss(self, phase=-1):
        self.lib.call("mmg_loss", C.byref(self.cfg), self.params.data_ptr(), C.byref(self._inp),
                      self.workspace.data_ptr(), int(phase), self._stream())

    def backward(self):
        self.lib.call("mmg_backward", C.byref(self.cfg), self.params.data_ptr(), C.byref(self._inp),
                      self.workspace.data_ptr(), self.grads.data_ptr(), self._stream())

    def grad_norm(self):
        self.lib.call("mmg_grad_norm", C.byref(self.cfg), self.grads.data_ptr(), self.workspace.data_ptr(), self._stream())

    def update(self):
        self.step += 1
        s2 = None if self.state2 is None else self.state2.data_ptr()
        self.lib.call("mmg_clip_update", C.byref(self.cfg), self.params.data_ptr(), self.grads.data_ptr(),
                      self.state1.data_ptr(), s2, C.c_int64(self.step), C.c_float(1.0), self.workspace.data_ptr(),
                      self._stream())

    def train_step(self, x, desc, target, uniforms=None, top_k=6):
        """One fused training iteration (model.py:1240-1339) on device-resident inputs."""
        self._inp = self._inputs(x, desc, target, True, uniforms, None, None, top_k)
        self.step += 1
        s2 = None if self.state2 is None else self.state2.data_ptr()
        self.lib.call("mmg_train_step", C.byref(self.cfg), self.params.data_ptr(), self.grads.data_ptr(),
                      self.state1.data_ptr(), s2, C.c_int64(self.step), C.byref(self._inp), self.workspace.data_ptr(),
                      self._stream())

    # ---- host-buffer pipeline (the end-to-end path: pinned host batches in, loss values out) -------------------------
    def enable_host_pipeline(self, desc, graphs=True):
        """Two device staging slots + a copy stream: the H2D copy of batch i+1 overlaps the training of batch i.
        `graphs`: capture the launch sequence of one iteration per slot in a CUDA graph and replay it (single GPU, RMSprop / SGD:
        nothing in the sequence changes from step to step; Adam's bias correction takes the step number as a kernel argument
        and the data-parallel paths use it as the flag value, so those keep the eager launches)."""
        d = self.dims
        dev = self.device
        self._hp = dict(
            copy_stream=torch.cuda.Stream(device=dev),
            x=[torch.empty(d["B"], d["F"], dtype=torch.float32, device=dev) for _ in range(2)],
            t=[torch.empty(d["B"], dtype=torch.int64, device=dev) for _ in range(2)],
            ready=[torch.cuda.Event() for _ in range(2)], free=[torch.cuda.Event() for _ in range(2)],
            desc=torch.as_tensor(desc).to(device=dev, dtype=torch.float32).contiguous(), used=[False, False], n=0)
        self._hp["inp"] = [self._inputs(self._hp["x"][i], self._hp["desc"], self._hp["t"][i], True, None, None, None, 6)
                           for i in range(2)]
        self._keep = None
        # everything that does not change from step to step is marshalled ONCE: the per-step Python work of the host
        # pipeline is two foreign calls with prebuilt arguments (at ~0.1 ms of device time per step the host must stay ahead)
        hp = self._hp
        cur = torch.cuda.current_stream(dev)
        for i in range(2):                       # materialise the event handles before the C calls re-record them
            hp["ready"][i].record(hp["copy_stream"]); hp["free"][i].record(cur)
        cs = C.c_void_p(hp["copy_stream"].cuda_stream)
        vp = lambda t: C.c_void_p(t.data_ptr())
        hp["pf_args"] = [(vp(hp["x"][i]), vp(hp["t"][i]), cs, C.c_void_p(hp["free"][i].cuda_event),
                          C.c_void_p(hp["ready"][i].cuda_event)) for i in range(2)]
        s2 = None if self.state2 is None else vp(self.state2)
        hp["ts_args"] = [(C.byref(self.cfg), vp(self.params), vp(self.grads), vp(self.state1), s2, C.byref(hp["inp"][i]),
                          vp(self.workspace), C.c_void_p(hp["ready"][i].cuda_event), C.c_void_p(hp["free"][i].cuda_event))
                         for i in range(2)]
        hp["cfg_ref"] = C.byref(self.cfg)
        hp["stream"] = cur
        hp["stream_ptr"] = C.c_void_p(cur.cuda_stream)
        hp["fn_pf"] = self.lib.dll.mmg_host_prefetch
        hp["fn_ts"] = self.lib.dll.mmg_train_step_staged
        hp["graphs"] = None
        hp["losses_dev"] = self.ws("losses", (capi.MMG_LOSS_COUNT,))
        # loss values of the iteration on staging slot i land here (mapped pinned memory written by the update kernel itself:
        # no device-to-host copy between two iterations); valid once that iteration has completed
        hp["h_losses"] = torch.zeros(2, capi.MMG_LOSS_COUNT, dtype=torch.float32).pin_memory()
        for i in range(2):
            hp["inp"][i].h_losses_out = C.c_void_p(hp["h_losses"][i].data_ptr())
        if (graphs and self.device.type == "cuda" and getattr(self, "_peers", None) is None
                and int(self.cfg.optim_type) != capi.OPTIM["Adam"] and os.environ.get("MMG_GRAPHS", "1") != "0"):
            try:
                gs = []
                for i in range(2):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        cfg, p, gr, s1, s2_, inp, ws, _, _ = hp["ts_args"][i]
                        self.lib.call("mmg_train_step", cfg, p, gr, s1, s2_, C.c_int64(1), inp, ws,
                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
                    gs.append(g)
                hp["graphs"] = gs
            except Exception:           # capture refused (driver / runtime restriction): the eager launch sequence does the same work
                hp["graphs"] = None
                torch.cuda.synchronize(dev)

    def host_prefetch(self, h_x, h_target):
        """Enqueue the H2D copy of the NEXT batch (pinned host tensors) into the free staging slot."""
        hp = self._hp
        slot = hp["n"] & 1
        dx, dt, cs, free, ready = hp["pf_args"][slot]
        rc = hp["fn_pf"](hp["cfg_ref"], h_x.data_ptr(), h_target.data_ptr(), dx, dt, cs, free if hp["used"][slot] else None, ready)
        if rc < 0:
            self.lib.check(rc, "mmg_host_prefetch")
        hp["n"] += 1
        hp["pending"] = slot

    def train_step_staged(self, h_losses, slot, dp_group=None):
        """Train on staging slot `slot` (its prefetch was enqueued earlier).  `h_losses` (pinned tensor) receives the loss values
        through a device-to-host copy; pass None to read them from `staged_losses(slot)` instead — pinned memory the update kernel
        writes itself, valid once the iteration has completed and until that slot trains again (no copy between iterations).
        Data-parallel engines (enable_peer_dp, or `dp_group` for the NCCL variant) run their own iteration on the slot."""
        hp = self._hp
        if getattr(self, "_peers", None) is not None or dp_group is not None:
            st = torch.cuda.current_stream(self.device)
            st.wait_event(hp["ready"][slot])
            if dp_group is not None:
                self._inp = hp["inp"][slot]
                self._dp_iteration(dp_group)
            else:
                self._peer_iteration(hp["inp"][slot])
            hp["free"][slot].record(st)
            hp["used"][slot] = True
            if h_losses is not None or dp_group is not None:       # the peer iteration delivers into staged_losses(slot) itself
                (h_losses if h_losses is not None else hp["h_losses"][slot]).copy_(hp["losses_dev"], non_blocking=True)
            return
        self.step += 1
        if hp["graphs"] is not None:
            st = hp["stream"]
            st.wait_event(hp["ready"][slot])
            hp["graphs"][slot].replay()
            hp["free"][slot].record(st)
            if h_losses is not None:
                h_losses.copy_(hp["losses_dev"], non_blocking=True)
            hp["used"][slot] = True
            return
        cfg, p, g, s1, s2, inp, ws, ready, free = hp["ts_args"][slot]
        rc = hp["fn_ts"](cfg, p, g, s1, s2, self.step, inp, ws, None if h_losses is None else h_losses.data_ptr(), hp["stream_ptr"],
                         ready, free)
        if rc < 0:
            self.lib.check(rc, "mmg_train_step_staged")
        hp["used"][slot] = True

    def staged_losses(self, slot):
        """Loss values (capi.LOSS_NAMES order) of the last iteration trained on staging slot `slot` (pinned host tensor)."""
        return self._hp["h_losses"][slot]

    # ---- data parallel over NVLink peer memory (no collective call, no extra launch) ---------------------------------
    def enable_peer_dp(self, group=None):
        """Allocate this rank's symmetric exchange buffer (torch symmetric memory: peer-mapped on every rank of `group`)
        and resolve the peers' pointers for `train_step_peer`.  Collective: every rank of the group must call it."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        tot, so, ro, sto, no, fo = [C.c_int64() for _ in range(6)]
        self.lib.call("mmg_peer_buffer_layout", C.byref(self.cfg), C.byref(tot), C.byref(so), C.byref(ro), C.byref(sto),
                      C.byref(no), C.byref(fo))
        buf = symm_mem.empty(int(tot.value), dtype=torch.uint8, device=self.device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, group)
        world, rank = int(hdl.world_size), int(hdl.rank)
        if world > capi.MMG_MAX_PEERS:
            raise capi.MmgError("peer data parallelism supports up to %d ranks" % capi.MMG_MAX_PEERS)
        p = capi.Peers()
        p.world, p.rank = world, rank
        for r in range(world):
            base = int(hdl.buffer_ptrs[r])
            p.d_send[r], p.d_recv[r], p.d_stats[r] = base + so.value, base + ro.value, base + sto.value
            p.d_norms[r], p.d_flags[r] = base + no.value, base + fo.value
        # in-switch reduction of the gradient slices when the symmetric buffer has a multicast mapping (NVSwitch, NVLS); MMG_NVLS=0:
        # one peer load per rank instead
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        p.d_send_mc = C.c_void_p(mc + so.value) if (mc and os.environ.get("MMG_NVLS", "1") != "0") else None
        self.peer_nvls = bool(p.d_send_mc)
        self._peer_err = torch.zeros(1, dtype=torch.int32, device=self.device)
        p.d_error = self._peer_err.data_ptr()
        self._peers, self._peer_buf, self._peer_hdl = p, buf, hdl
        self.peer_check_every = 1024
        torch.cuda.synchronize(self.device)
        hdl.barrier()                      # every rank's buffer is zeroed before anyone's first flag lands
        torch.cuda.synchronize(self.device)
        return world, rank

    def train_step_peer(self, x, desc, target, uniforms=None, top_k=6):
        """Data-parallel iteration: this rank's batch shard; batch statistics and gradients are summed across the ranks
        inside the kernels over NVLink peer memory (mmg_train_step_peer)."""
        self._inp = self._inputs(x, desc, target, True, uniforms, None, None, top_k)
        self._peer_iteration(self._inp)

    def _peer_iteration(self, inp):
        # a timed-out peer wait is sticky on the device (updates are skipped from then on): surface it regularly
        if self.step and self.step % self.peer_check_every == 0 and self.peer_error():
            raise capi.MmgError("a peer wait timed out (error %d): the ranks lost lockstep; parameters were left untouched "
                                "from that iteration on" % self.peer_error())
        self.step += 1
        s2 = None if self.state2 is None else self.state2.data_ptr()
        self.lib.call("mmg_train_step_peer", C.byref(self.cfg), self.params.data_ptr(), self.grads.data_ptr(),
                      self.state1.data_ptr(), s2, C.c_int64(self.step), C.byref(inp), self.workspace.data_ptr(),
                      C.byref(self._peers), self._stream())

    def peer_error(self):
        """Non-zero when a peer wait timed out (synchronises)."""
        return int(self._peer_err.item())

    def train_step_dp(self, x, desc, target, group=None, uniforms=None, top_k=6):
        """Data-parallel iteration: this rank's batch shard; batch statistics and gradients are all-reduced
        (SURVEY.md §8e).  Two collectives per iteration: a few hundred doubles, then the flat gradient buffer."""
        self._inp = self._inputs(x, desc, target, True, uniforms, None, None, top_k)
        self._dp_iteration(group)

    def _dp_iteration(self, group):
        import torch.distributed as dist
        self.lib.call("mmg_exchange_forward", C.byref(self.cfg), self.params.data_ptr(), C.byref(self._inp),
                      self.workspace.data_ptr(), self._stream())
        self.loss(0)
        dist.all_reduce(self.stats(), group=group)
        self.loss(1)
        self.backward()
        dist.all_reduce(self.grads, group=group)
        self.grad_norm()
        self.update()

    # ---- results -----------------------------------------------------------------------------------------------
    def stats(self):
        return self.ws("stats", (int(self.wsl.stats_count),), torch.float64)

    def losses(self):
        """dict of loss values (device->host copy; synchronises)."""
        v = self.ws("losses", (capi.MMG_LOSS_COUNT,)).detach().cpu().tolist()
        return dict(zip(capi.LOSS_NAMES, v))

    def outputs(self):
        """Views of everything `exchange()` returns (model.py:872-876), stacked over steps."""
        d = self.dims
        T, B, M, D, Hr, Hi = d["T"], d["B"], d["M"], d["D"], d["Hr"], d["Hi"]
        return dict(
            sen_feats=self.ws("sen_feats", (T, B, M)), sen_probs=self.ws("sen_probs", (T, B, M)),
            rec_feats=self.ws("rec_feats", (T + 1, B, M))[1:], rec_probs=self.ws("rec_probs", (T, B, M)),
            stop_feat=self.ws("stop_feat", (T, B, 1)), stop_prob=self.ws("stop_prob", (T, B, 1)),
            y=self.ws("y", (T, B, D)), stop_mask=self.ws("stop_mask", (T + 1, B, 1), torch.uint8),
            bs=self.ws("bs", (T, B, 1)), br=self.ws("br", (T, B, 1)), h_x=self.ws("h_x", (B, Hi)),
            h_z=self.ws("h_z", (T + 1, B, Hr))[1:], h_w=self.ws("h_w", (T, B, Hr)),
            outp=self.ws("outp", (B, D)), logs=self.ws("logs", (B, 1)), argmax=self.ws("argmax", (B,), torch.int32),
            ystep=self.ws("ystep", (B,), torch.int32), grad_norms=self.ws("grad_norms", (4,)))
