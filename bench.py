#!/usr/bin/env python
"""bench.py — exchange-steps/sec of the referential-game training iteration (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4|C5|C2A]

One "step" = one full training iteration (model.py:1240-1339: T-step conversation, losses, backward, 4x clip +
RMSprop) over one synthetic batch; the metric is exchange steps per second = T' x iterations / second, whole job.
N=1 runs configs[1] of BASELINE.json (the configuration the metric is quoted on): fixed 10-step exchange, batch 64,
30 classes, 2048-d features, -use_binary.  N>1 (torchrun, one rank per GPU) keeps 64 rows per GPU (weak scaling,
configs[3] at N=8); the batch statistics and the flat gradient are summed across the ranks inside the kernels over NVLink
peer memory (or by NCCL with --dp nccl).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU port of the reference path (oracle/game_oracle.py) on
the host cores, on the SAME global batch, honouring --steps / --warmup (the reference itself is Python 2 / torch 0.1.12
code that only runs under the build container's compat shim).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

HEAD = dict(img_h_dim=256, baseline_hid_dim=500, sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100,
            entropy_sen=0.01, entropy_rec=0.01, top_k_train=6)
CONFIGS = {
    "C2": dict(batch_size=64, img_feat_dim=2048, n_classes=30, max_exchange=10, fixed_exchange=True, use_binary=True, **HEAD),
    "C3": dict(batch_size=256, img_feat_dim=2048, n_classes=30, max_exchange=10, fixed_exchange=False, use_binary=True,
               entropy_s=0.08, **HEAD),
    "C4": dict(batch_size=64, img_feat_dim=2048, n_classes=30, max_exchange=10, fixed_exchange=True, use_binary=True, **HEAD),
    "C5": dict(batch_size=128, img_feat_dim=2048, n_classes=100, max_exchange=20, fixed_exchange=False, use_binary=True,
               entropy_s=0.08, **dict(HEAD, sender_out_dim=64, rec_w_dim=64)),
    # C2 with the receiver's word-level description attention (-desc_attn, model.py:344-410): 30 classes x 3..14 words
    "C2A": dict(batch_size=64, img_feat_dim=2048, n_classes=30, max_exchange=10, fixed_exchange=True, use_binary=True,
                desc_attn=True, desc_attn_dim=64, **HEAD),
}
WORKLOAD = {
    "C2": "BASELINE.json configs[1]: fixed 10-step exchange, batch=64, 30 classes, 2048-d feats, -use_binary",
    "C3": "BASELINE.json configs[2]: adaptive max_exchange=10, batch=256, 30 classes, entropy_s=0.08",
    "C4": "BASELINE.json configs[3]: fixed 10-step, 64 rows per GPU, 30 classes, -use_binary, data-parallel",
    "C5": "BASELINE.json configs[4]: adaptive max_exchange=20, 128 rows per GPU, 100 classes, rec_w_dim=64",
    "C2A": "configs[1] with -desc_attn -desc_attn_dim 64: fixed 10-step, batch=64, 30 classes x 3..14 description words",
}


def bench_config(name, cfgd, world):
    """The `config` object of the JSON line: identical for both arms (`--impl ours` and `--impl reference`)."""
    return {"workload": WORKLOAD[name], "name": name, "batch_per_gpu": cfgd["batch_size"],
            "global_batch": cfgd["batch_size"] * world, "max_exchange": cfgd["max_exchange"],
            "fixed_exchange": bool(cfgd["fixed_exchange"]), "n_classes": cfgd["n_classes"], "img_feat_dim": cfgd["img_feat_dim"],
            "msg_dim": cfgd["rec_w_dim"], "parallelism": "dp%d" % world}


def param_counts(c):
    F, Hi, M, Hr, WV, Hb = c["img_feat_dim"], c["img_h_dim"], c["rec_w_dim"], c["rec_hidden"], c["wv_dim"], c["baseline_hid_dim"]
    sender = Hi * F + Hi + Hi * M + Hi + M + M * Hi + M
    receiver = 3 * Hr * M + 3 * Hr * Hr + 6 * Hr + Hr * Hr + Hr + Hr * WV + M * Hr + M + Hr * (Hr + WV) + Hr + Hr + 1 + Hr + 1
    bas = Hb * (Hi + M) + Hb + Hb + 1 + Hb * (M + Hr) + Hb + Hb + 1
    if c.get("desc_attn"):
        A = c["desc_attn_dim"]
        receiver += A * WV + A + A * Hr + A + A + 1
    return sender + receiver + bas


def algorithmic_bytes_per_iteration(c, n_gpus):
    """SURVEY.md §8(d) / BASELINE.md §4.7: inputs once, returned per-step outputs, forward read of every parameter,
    gradient write + optimizer read/write of parameter and state (+ all-reduce send/recv when data-parallel)."""
    B, F, D, WV, T, M = c["batch_size"], c["img_feat_dim"], c["n_classes"], c["wv_dim"], c["max_exchange"], c["rec_w_dim"]
    P = param_counts(c)
    b = 4 * (B * F + D * WV) + 8 * B + 4 * T * B * (4 * M + D + 4) + 4 * P + 20 * P
    if n_gpus > 1:
        b += 8 * P
    return b


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the run (B200_PROFILING.md recipe); every sample carries its
    arrival time so that only samples taken under load (between `mark_begin` and `mark_end`) are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.t0, self.t1 = None, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [f.strip() for f in line.split(",")]))
        except Exception:
            pass

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if self.t0 is not None and not (self.t0 <= ts <= (self.t1 or ts)):
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "window": "untimed pre-conditioning burst of the same step (>= 0.6 s) + warm-up + timed region"}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference path's port (oracle/game_oracle.py), full training iterations incl. host-side sampling
# ---------------------------------------------------------------------------------------------------------------------
class CpuArm(object):
    def __init__(self, cfgd, world):
        from oracle import game_oracle as go
        from multimodalgame_b200 import synthetic as syn
        self.go = go
        gd = dict(cfgd, batch_size=cfgd["batch_size"] * world)          # the SAME workload: the global batch
        self.cfg = go.GameConfig(**gd)
        fl = syn.GameFlags(**gd)
        self.params = go.clone_params(syn.init_params(fl, seed=0))
        self.state = go.new_opt_state(self.params)
        self.x, self.desc, self.target = syn.batch(fl, seed=0)
        self.words = syn.desc_set(fl, seed=0)
        self.rng = np.random.RandomState(0)

    def iteration(self):
        us = self.go.draw_uniforms(self.rng, self.cfg)
        t0 = time.perf_counter()
        ex, _ = self.go.train_iteration(self.params, self.state, self.x, self.target, self.desc, self.cfg, us, **self.words)
        return time.perf_counter() - t0, len(ex["y"])

    def run(self, iters, warmup, threads):
        torch.set_num_threads(threads)
        for _ in range(warmup):
            self.iteration()
        total, steps = 0.0, 0
        for _ in range(iters):
            dt, n = self.iteration()
            total += dt; steps += n
        return total, steps


def cpu_baseline_sample(cfgd, world, budget_s=12.0, max_iters=30):
    """Bounded sample (about 10-30 s of CPU work) of the same workload with all host threads AND one thread; best reported."""
    arm = CpuArm(cfgd, world)
    ncpu = os.cpu_count() or 1
    best = None
    for threads in sorted({ncpu, 1}, reverse=True):
        torch.set_num_threads(threads)
        dt, _ = arm.iteration()                                         # warm-up + cost estimate
        iters = int(max(2, min(max_iters, budget_s / max(dt, 1e-4))))
        total, steps = arm.run(iters, 1, threads)
        r = dict(value=world * steps / total, threads=threads, iters=iters, ms=1e3 * total / iters)
        if best is None or r["value"] > best["value"]:
            best = r
    return best, ncpu


def run_reference(args, cfgd, name, out_stream):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    ncpu = os.cpu_count() or 1
    arm = CpuArm(cfgd, world)
    K, Wm = max(1, args.steps if args.steps else 20), max(0, args.warmup)
    torch.set_num_threads(ncpu)            # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core it can
    t_probe, _ = arm.iteration()
    threads = ncpu
    if ncpu > 1:                           # the tiny per-op GEMMs of this path do not always scale with threads: probe one thread
        torch.set_num_threads(1)
        arm.iteration()
        t1, _ = arm.iteration()
        torch.set_num_threads(ncpu)
        tn, _ = arm.iteration()
        if t1 < tn:
            threads, t_probe = 1, t1
        else:
            t_probe = tn
    n_timed = K if K * t_probe < 150.0 else max(1, int(150.0 / t_probe))       # bounded: the whole run ends within minutes
    total, steps = arm.run(n_timed, Wm, threads)
    val = world * steps / total            # same unit as the GPU arm: one exchange step over one batch shard (N shards per global step)
    sample = ("%d training iterations of the same workload (global batch %d = %d shard(s) of %d rows, each exchange step counted "
              "once per shard like the GPU arm) on the host: oracle/game_oracle.py, torch CPU fp32, %d threads (os.cpu_count=%d; "
              "all-threads vs 1-thread probed, faster one used), %.1f ms/iteration"
              % (n_timed, arm.cfg.batch_size, world, cfgd["batch_size"], threads, ncpu, 1e3 * total / n_timed))
    out = {"impl": "reference", "metric": "exchange-steps/sec", "value": val, "unit": "exchange-steps/s", "n_gpus": args.gpus,
           "steps": n_timed, "warmup": Wm, "ms_per_step": 1e3 * total / n_timed, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": bench_config(name, cfgd, world),
           "cpu_baseline": {"value": val, "unit": "exchange-steps/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "exchange-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out_stream.write(json.dumps(out) + "\n")
    out_stream.flush()


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for one), so
    everything written to fd 1 during the run goes to stderr and the JSON line is written to the original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def dp_oracle_check(e, fl, cfgd, world, rank, dev, dist, step_on):
    """One data-parallel iteration of the timed configuration (on-device sampler) against the CPU oracle on the GLOBAL batch:
    the bits every rank drew are gathered and replayed through the oracle (u = 1 - bit reproduces the bit for any
    probability); class scores, losses and post-step parameters must agree.  Rank 0 computes the oracle and raises on a
    parity violation; returns a summary dict."""
    from oracle import game_oracle as go
    from multimodalgame_b200 import synthetic as syn
    B, T = fl.batch_size, fl.max_exchange

    def close(what, got, want, tol=1e-4):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        err = float(np.max(np.abs(got - want) - tol * np.abs(want))) if got.size else 0.0
        assert err <= tol, "data-parallel check: %s off by %.3g (> %.0e)" % (what, err + tol, tol)
        return float(np.max(np.abs(got - want))) if got.size else 0.0

    gcfg = go.GameConfig(**dict(cfgd, batch_size=B * world))
    gfl = syn.GameFlags(**dict(cfgd, batch_size=B * world))
    params0 = {a: {k: v.detach().cpu().clone() for k, v in d_.items()} for a, d_ in e.named_views().items()}
    words = syn.desc_set(fl, seed=0)
    x, desc, target = syn.batch(gfl, seed=507)
    sl = slice(rank * B, (rank + 1) * B)
    step_on(x[sl].to(dev), desc.to(dev), target[sl].to(dev))
    torch.cuda.synchronize(dev)
    o = e.outputs()
    from multimodalgame_b200 import capi
    lv = e.ws("losses", (capi.MMG_LOSS_COUNT,)).double().clone()
    dist.all_reduce(lv)                        # every rank reports its contribution to the global means
    L = dict(zip(capi.LOSS_NAMES, lv.cpu().tolist()))
    gathered = {}
    for key in ("sen_feats", "rec_feats", "stop_feat", "y"):
        t = o[key].contiguous()
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        gathered[key] = torch.cat([p_.cpu() for p_ in parts], 1).numpy()
    summary = {}
    if rank == 0:
        G = B * world
        us = [(1.0 - gathered["sen_feats"][t].astype(np.float64), 1.0 - gathered["stop_feat"][t].astype(np.float64).reshape(G, 1),
               1.0 - gathered["rec_feats"][t].astype(np.float64)) for t in range(T)]
        oparams = go.clone_params(params0)
        ex, res = go.train_iteration(oparams, go.new_opt_state(oparams), x, target, desc, gcfg, us, **words)
        Tp = len(ex["y"])
        st = lambda key: np.stack([t_.detach().numpy() for t_ in ex[key]], 0)
        assert np.array_equal(gathered["sen_feats"][:Tp], st("sen_feats")), "data-parallel check: replayed sender bits differ"
        assert not np.array_equal(gathered["sen_feats"][:, :B], gathered["sen_feats"][:, B:2 * B]), "ranks drew identical noise"
        summary["y_max_err"] = close("y", gathered["y"][:Tp], st("y"))
        for nm in ("nll_loss", "loss_rec", "loss_sen", "loss_bas_rec", "loss_bas_sen"):
            summary[nm + "_err"] = close(nm, L[nm], float(res[nm].detach()))
        pv = e.named_views()
        lr = fl.learning_rate
        worst = 0.0
        for a in oparams:
            for k, v in oparams[a].items():
                if (a, k) in (("receiver", "y2.bias"), ("receiver", "d_attn.bias")):
                    continue
                worst = max(worst, float((pv[a][k].detach().cpu() - v).abs().max()))
        assert worst <= 12 * lr, "data-parallel check: post-step parameters off by %.3g (> 12 lr)" % worst
        summary["param_max_err_over_lr"] = worst / lr
        summary["global_batch"] = G
        summary["what"] = "1 iteration, on-device sampler, bits replayed through oracle/game_oracle.py on the global batch"
    return summary


def main():
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed iterations (default: as many as fill ~0.6 s)")
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the secondary measurement that flushes L2 between steps")
    ap.add_argument("--ring-mib", type=int, default=160, help="size of the ring of distinct input batches (must exceed L2)")
    ap.add_argument("--no-dp-check", action="store_true", help="N>1: skip the global-batch oracle check before timing")
    ap.add_argument("--dp", default="peer", choices=["peer", "nccl"],
                    help="N>1: in-kernel NVLink peer-memory reduction (default) or torch.distributed NCCL all-reduce")
    args = ap.parse_args()
    name = args.config or ("C2" if args.gpus == 1 else "C4")
    cfgd = dict(CONFIGS[name])
    if args.impl == "reference":
        run_reference(args, cfgd, name, out_stream)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    sampler = ClockSampler(local)
    sampler.start()                       # nvidia-smi takes a few hundred ms to produce its first line: start it first

    import __graft_entry__ as ge
    ge.build()
    from multimodalgame_b200 import capi, engine as eng, synthetic as syn

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = capi.load()
    fl = syn.GameFlags(**cfgd)
    B, T = fl.batch_size, fl.max_exchange
    words = syn.desc_set(fl, seed=0)
    # every rank uses the SAME sampler seed: the Philox counters are keyed by the global row (batch_offset)
    e = eng.GameEngine(syn.config_from_flags(fl, batch_global=B * world, n_words=int(words["desc_set"].shape[0]) if words else 0,
                                             batch_offset=rank * B), device=dev, lib=lib, seed=1)
    e.load_params(syn.init_params(fl, seed=0))
    if words:
        e.set_desc_set(**words)
    # synthetic inputs: a ring of distinct batches LARGER THAN L2 (126 MB), resident in HBM for `value`, in pinned host memory
    # for `e2e`: every timed step reads a batch that cannot be cache resident, while parameters / optimizer state / workspace
    # stay wherever the previous step left them, as in a real training loop
    x_bytes = B * fl.img_feat_dim * 4
    nb = max(8, -(-args.ring_mib * (1 << 20) // x_bytes))
    gen = torch.Generator().manual_seed(1000 + rank)
    desc = syn.batch(fl, seed=0)[1].to(dev)             # class descriptions are shared by all ranks
    hx_all = torch.randn(nb, B, fl.img_feat_dim, generator=gen).pin_memory()           # x ~ N(0,1), SURVEY.md 8(d)
    ht_all = torch.randint(0, fl.n_classes, (nb, B), generator=gen, dtype=torch.int64).pin_memory()
    xs_all, ts_all = hx_all.to(dev), ht_all.to(dev)
    xs, ts = list(xs_all.unbind(0)), list(ts_all.unbind(0))
    hx, ht = list(hx_all.unbind(0)), list(ht_all.unbind(0))
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    dp_mode = "none"
    if world > 1:
        dp_mode = "nccl"
        if args.dp == "peer":
            try:
                e.enable_peer_dp()
                dp_mode = "peer"
            except Exception as ex:      # symmetric memory unavailable on this box: fall back to the NCCL path, and say so
                sys.stderr.write("peer data-parallel path unavailable (%s); using NCCL all-reduce\n" % (ex,))
        flag = torch.tensor([1 if dp_mode == "peer" else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 0:
            dp_mode = "nccl"

    def step_on(x, d, t, uniforms=None):
        if dp_mode == "peer":
            e.train_step_peer(x, d, t, uniforms=uniforms)
        elif dp_mode == "nccl":
            e.train_step_dp(x, d, t, uniforms=uniforms)
        else:
            e.train_step(x, d, t, uniforms=uniforms)

    def step(i):
        step_on(xs[i % nb], desc, ts[i % nb])

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- N>1: parity of the data-parallel iteration with the global-batch oracle, before anything is timed ------------
    dp_check = None
    if world > 1 and not args.no_dp_check:
        dp_check = dp_oracle_check(e, fl, cfgd, world, rank, dev, dist, step_on)
        sync_all()

    # ---- untimed pre-conditioning burst (clocks ramp up, sampler collects under-load samples), then W warm-up steps -----
    sampler.mark_begin()
    t_burst = time.perf_counter()
    n_burst = 0
    while True:
        for i in range(50):
            step(n_burst + i)
        n_burst += 50
        torch.cuda.synchronize(dev)
        go_on = torch.tensor([1 if time.perf_counter() - t_burst < 0.6 else 0], device=dev)
        if dist is not None:
            dist.all_reduce(go_on, op=dist.ReduceOp.MAX)       # every rank runs the same number of iterations
        if int(go_on) == 0:
            break
    burst_ms_per_iter = 1e3 * (time.perf_counter() - t_burst) / n_burst
    W = max(3, args.warmup)
    for i in range(W):
        step(i)
    sync_all()
    if args.steps:
        K = args.steps
    else:                                  # default: a timed region of >= 0.6 s (a handful of nvidia-smi samples fit inside)
        kk = torch.tensor([int(min(20000, max(200, math.ceil(600.0 / max(burst_ms_per_iter, 1e-3)))))], device=dev)
        if dist is not None:
            dist.all_reduce(kk, op=dist.ReduceOp.MAX)
        K = int(kk)
    # ---- primary measurement: EXACTLY K steps back to back over the input ring (inputs larger than L2), one pair of CUDA
    #      events on the launching stream, barrier + synchronize on both sides
    lib.dll.mmg_launch_count_reset()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall = time.perf_counter()
    a0.record()
    for i in range(K):
        step(W + i)
    a1.record()
    sync_all()
    t_wall = time.perf_counter() - t_wall
    launches = lib.dll.mmg_launch_count()
    tot_ms = a0.elapsed_time(a1)
    active = e.losses()["active_steps"]
    # ---- secondary: the round-1 protocol, L2 flushed by a 256 MiB memset before every step (outside the per-step events);
    #      the memset leaves L2 full of dirty lines, so this also charges their write-back to the step
    flushed_ms = None
    if flush is not None:
        Kf = min(K, 400)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kf)]
        sync_all()
        for i in range(Kf):
            flush.zero_()
            ev[i][0].record()
            step(i)
            ev[i][1].record()
        sync_all()
        flushed_ms = float(sum(a.elapsed_time(b) for a, b in ev)) / Kf
    sampler.mark_end()
    clocks = sampler.stop()
    steps_per_iter = T if fl.fixed_exchange else float(active)

    # ---- e2e: host buffers through the public API: every step copies its batch from pinned host memory (H2D, on a copy
    #      stream, double-buffered: overlaps the previous step) and its loss values back (D2H).  Timed on the device with
    #      one pair of events around the K steps (plus the host wall clock), max over ranks.
    e.enable_host_pipeline(desc)

    def run_host(n):
        e.host_prefetch(hx[0], ht[0])
        for i in range(n):
            slot = i % 2
            if i + 1 < n:
                e.host_prefetch(hx[(i + 1) % nb], ht[(i + 1) % nb])
            # loss values -> pinned host memory every step: written by the update kernel itself into the slot's mapped pinned
            # buffer (single GPU), or copied device-to-host behind the step (data-parallel paths)
            e.train_step_staged(None, slot, dp_group=dist.group.WORLD if dp_mode == "nccl" else None)
    run_host(6)
    sync_all()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    h0.record()
    run_host(K)
    h1.record()
    e2e_enqueue_ms = 1e3 * (time.perf_counter() - t0)         # host time to enqueue the K steps (before any synchronisation)
    torch.cuda.synchronize(dev)
    e2e_wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max(h0.elapsed_time(h1), 0.0)
    last_losses = e.staged_losses((K - 1) % 2).clone()
    assert float(last_losses[0]) != 0.0 and bool(torch.isfinite(last_losses).all())     # the loss values really arrived on the host
    dev_losses = e.ws("losses", (capi.MMG_LOSS_COUNT,)).cpu()
    assert torch.equal(last_losses, dev_losses), "host copy of the loss values differs from the device's"

    # ---- N>1: the replicas must still be bit-identical and no peer wait may have timed out ----------------------------------
    replicas = None
    t = torch.tensor([tot_ms, flushed_ms if flushed_ms is not None else 0.0, e2e_ms, e2e_wall_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if dp_mode == "peer":
            assert e.peer_error() == 0, "a peer wait timed out on rank %d (error %d): the numbers are void" % (rank, e.peer_error())
        digest = torch.stack([e.params.view(torch.int32).to(torch.int64).sum(), (e.params.double() ** 2).sum().view(torch.int64)])
        alld = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(alld, digest)
        same = all(bool(torch.equal(alld[0], d_)) for d_ in alld)
        assert same, "parameter replicas diverged across ranks: %s" % [d_.tolist() for d_ in alld]
        replicas = "bit-identical on %d ranks (int32 checksum + sum of squares of the flat parameter buffer)" % world
    tot_ms, flushed_ms_max, e2e_ms, e2e_wall_ms = [float(v) for v in t]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    e2e = {"value": world * steps_per_iter * K / (e2e_ms * 1e-3), "unit": "exchange-steps/s",
           "h2d_bytes_per_step": int(world * (B * fl.img_feat_dim * 4 + B * 8)),
           "d2h_bytes_per_step": int(world * capi.MMG_LOSS_COUNT * 4), "ms_per_step": e2e_ms / K,
           "ms_per_step_host_wall": e2e_wall_ms / K, "ms_per_step_host_enqueue": e2e_enqueue_ms / K,
           "launch": "CUDA graph replay per staging slot" if getattr(e, "_hp", {}).get("graphs") else "eager launches",
           "note": "mmg_host_prefetch + mmg_train_step_staged (N>1: the same staging slots feeding mmg_train_step_peer): pinned "
                   "host x/target -> device on a copy stream (double-buffered, overlaps the previous step), losses -> pinned "
                   "host every step (single GPU: written by the update kernel into mapped pinned memory; N>1: async copy); bytes are the whole job's (all ranks); device-timed, max over ranks"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bytes_iter = algorithmic_bytes_per_iteration(cfgd, world)
    # DRAM traffic per iteration from the committed ncu capture of this same command (profiles/)
    traffic, traffic_src, traffic_protocol = None, None, None
    if name in ("C2", "C2A") and world == 1:
        try:
            import glob
            pat = "*_dram_traffic.json" if name == "C2" else "*_c2a_desc_attn_dram.json"
            f = sorted(glob.glob(os.path.join(ROOT, "profiles", pat)))[-1]
            tj = json.load(open(f))
            traffic = float(tj["per_iteration"]["total_bytes"])
            traffic_src = os.path.relpath(f, ROOT)
            traffic_protocol = tj.get("protocol", "ncu default cache control: L2 flushed before EVERY kernel (per-kernel cold-cache "
                                                  "upper bound of the in-iteration traffic)")
        except Exception:
            pass
    ms_per_step = tot_ms / K
    achieved = bytes_iter / (ms_per_step * 1e-3) / 1e9
    value = world * steps_per_iter * K / (tot_ms * 1e-3)
    cfg_out = bench_config(name, cfgd, world)
    out = {
        "metric": "exchange-steps/sec", "value": value, "unit": "exchange-steps/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_out,
        "run": {"exchange_steps": steps_per_iter,
                "sampling": "on-device Philox4x32-10 keyed by the global batch row (parity tests inject the reference's float64 "
                            "uniforms instead)",
                "dp_reduction": {"peer": "in-kernel sums over NVLink peer memory (statistics + gradient); no collective call",
                                 "nccl": "torch.distributed NCCL all-reduce (statistics + flat gradient)",
                                 "none": "single GPU"}[dp_mode],
                "l2": "inputs larger than L2: ring of %d distinct batches (%.0f MiB of features per rank) cycled through the timed "
                      "steps, no flush; parameters, optimizer state and workspace stay cache resident as in a training loop"
                      % (nb, nb * x_bytes / float(1 << 20)),
                "preconditioning_iterations": n_burst, "dp_oracle_check": dp_check, "replicas": replicas},
        "value_l2_flushed": (world * steps_per_iter / (flushed_ms_max * 1e-3)) if flush is not None else None,
        "ms_per_step_l2_flushed": flushed_ms_max if flush is not None else None,
        "l2_flushed_protocol": "secondary, round-1 protocol: 256 MiB memset before every step, per-step CUDA events (the memset's dirty "
                               "lines are written back while the step runs)",
        "gpu_launches": int(launches), "launches_per_step": launches / float(K),
        "wall_s_timed_region": t_wall, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "traffic_protocol": traffic_protocol,
                     "peak_source": peak_src, "algorithmic_bytes_per_iteration": bytes_iter,
                     "kernel": "whole iteration = one launch sequence of %.0f kernels; achieved = algorithmic bytes per iteration / "
                               "CUDA-event time per iteration; the path is latency/dependency-bound, see DESIGN.md" % (launches / float(K))},
        "e2e": e2e,
    }
    if dist is not None:
        dist.destroy_process_group()
    if not args.no_cpu_baseline:
        best, ncpu = cpu_baseline_sample(cfgd, world)
        out["cpu_baseline"] = {"value": best["value"], "unit": "exchange-steps/s", "cores": best["threads"], "kind": "port",
                               "sample": "%d training iterations of the same workload (global batch %d; oracle/game_oracle.py, torch CPU "
                                         "fp32), %.1f ms/iteration with %d threads; all-threads (%d) and 1-thread both timed, best "
                                         "reported" % (best["iters"], B * world, best["ms"], best["threads"], ncpu)}
    out_stream.write(json.dumps(out) + "\n")
    out_stream.flush()


if __name__ == "__main__":
    main()
