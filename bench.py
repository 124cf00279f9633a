#!/usr/bin/env python
"""bench.py — exchange-steps/sec of the referential-game training iteration (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4|C5]

One "step" = one full training iteration (model.py:1240-1339: T-step conversation, losses, backward, 4x clip +
RMSprop) over one synthetic batch; the metric is exchange steps per second = T' x iterations / second, whole job.
N=1 runs configs[1] of BASELINE.json (the configuration the metric is quoted on): fixed 10-step exchange, batch 64,
30 classes, 2048-d features, -use_binary.  N>1 (torchrun, one rank per GPU) keeps 64 rows per GPU (weak scaling,
configs[3] at N=8) and all-reduces the batch statistics and the flat gradient buffer over NCCL.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference path on the host cores
(the reference itself is Python 2 / torch 0.1.12 code that only runs under the build container's compat shim).
"""
import argparse
import json
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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            for line in self.proc.stdout:
                self.rows.append([f.strip() for f in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def time_oracle(cfgd, iters, warmup, threads=None):
    """The reference path's CPU port (oracle/game_oracle.py): full training iterations incl. host-side sampling."""
    from oracle import game_oracle as go
    if threads:
        torch.set_num_threads(threads)
    cfg = go.GameConfig(**cfgd)
    params = go.init_params(cfg, seed=0)
    state = go.new_opt_state(params)
    x, desc, target = go.synthetic_batch(cfg, seed=0)
    from tests import parity_util as pu
    words = pu._synth_words(cfg, 0)
    rng = np.random.RandomState(0)
    times, steps = [], 0
    for i in range(warmup + iters):
        us = go.draw_uniforms(rng, cfg)
        t0 = time.perf_counter()
        ex, _ = go.train_iteration(params, state, x, target, desc, cfg, us, **words)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            steps += len(ex["y"])
    return sum(times), steps, torch.get_num_threads()


def run_reference(args, cfgd, name, out_stream):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    iters = max(1, min(args.steps, 40))
    total, steps, threads = time_oracle(cfgd, iters, min(args.warmup, 3))
    val = steps / total
    out = {"impl": "reference", "metric": "exchange-steps/sec", "value": val, "unit": "exchange-steps/s", "n_gpus": args.gpus,
           "steps": iters, "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * total / iters, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD[name], "batch": cfgd["batch_size"], "parallelism": "cpu"},
           "cpu_baseline": {"value": val, "unit": "exchange-steps/s", "cores": threads, "kind": "port",
                            "sample": "%d training iterations of the same workload on the host (oracle/game_oracle.py, "
                                      "torch CPU fp32, %d threads, os.cpu_count=%d)" % (iters, threads, os.cpu_count())},
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


def main():
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--dp", default="peer", choices=["peer", "nccl"],
                    help="N>1: in-kernel NVLink peer-memory reduction (default) or torch.distributed NCCL all-reduce")
    args = ap.parse_args()
    name = args.config or ("C2" if args.gpus == 1 else "C4")
    cfgd = dict(CONFIGS[name])
    if args.impl == "reference":
        run_reference(args, cfgd, name, out_stream)
        return

    import __graft_entry__ as ge
    ge.build()
    from multimodalgame_b200 import capi, engine as eng
    from oracle import game_oracle as go     # cpu_baseline leg + synthetic inputs only
    from tests import parity_util as pu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = capi.load()
    cfg = go.GameConfig(**cfgd)
    B, T = cfg.batch_size, cfg.max_exchange
    words = pu._synth_words(cfg, 0)
    e = eng.GameEngine(pu.config_from(cfg, batch_global=B * world, n_words=int(words["desc_set"].shape[0]) if words else 0),
                       device=dev, lib=lib, seed=1 + rank * 0)
    e.load_params(go.init_params(cfg, seed=0))
    if words:
        e.set_desc_set(**words)
    # synthetic inputs: a ring of distinct batches, resident in HBM for `value`, in pinned host memory for `e2e`
    nb = 8
    batches = [go.synthetic_batch(cfg, seed=100 * rank + i) for i in range(nb)]
    desc = batches[0][1].to(dev)
    xs = [b[0].to(dev) for b in batches]
    ts = [b[2].to(dev) for b in batches]
    hx = [b[0].pin_memory() for b in batches]
    ht = [b[2].pin_memory() for b in batches]
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

    def step(i):
        if dp_mode == "peer":
            e.train_step_peer(xs[i % nb], desc, ts[i % nb])
        elif dp_mode == "nccl":
            e.train_step_dp(xs[i % nb], desc, ts[i % nb])
        else:
            e.train_step(xs[i % nb], desc, ts[i % nb])

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(max(3, args.warmup)):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    lib.dll.mmg_launch_count_reset()
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sync_all()
    t_wall = time.perf_counter()
    for i in range(K):
        if flush is not None:
            flush.zero_()                      # evict L2 (126 MB) between timed steps; outside the per-step events
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    sync_all()
    t_wall = time.perf_counter() - t_wall
    launches = lib.dll.mmg_launch_count()
    ms = [a.elapsed_time(b) for a, b in ev]
    tot_ms = float(sum(ms))
    active = e.losses()["active_steps"]
    # back-to-back (L2-warm) run of the same K steps, one pair of events
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    a0.record()
    for i in range(K):
        step(i)
    a1.record()
    sync_all()
    warm_ms = a0.elapsed_time(a1)
    clocks = sampler.stop()
    t = torch.tensor([tot_ms, warm_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_ms, warm_ms = float(t[0]), float(t[1])
    steps_per_iter = T if cfg.fixed_exchange else float(active)

    # ---- e2e: host buffers through the C-ABI (mmg_host_prefetch + mmg_train_step_staged): every step copies its batch
    #      from pinned host memory (H2D, on a copy stream, overlapping the previous step) and its loss values back (D2H)
    e2e = None
    if world == 1:
        hl = torch.zeros(K, capi.MMG_LOSS_COUNT, dtype=torch.float32).pin_memory()
        e.enable_host_pipeline(desc)

        def run_host(n):
            e.host_prefetch(hx[0], ht[0])
            for i in range(n):
                slot = i % 2
                if i + 1 < n:
                    e.host_prefetch(hx[(i + 1) % nb], ht[(i + 1) % nb])
                e.train_step_staged(hl[i % K], slot)
        run_host(6)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        run_host(K)
        torch.cuda.synchronize(dev)
        e2e_wall = time.perf_counter() - t0
        e2e_ms = 1e3 * e2e_wall
        assert float(hl[K - 1][0]) != 0.0      # the loss values really arrived on the host
        e2e = {"value": steps_per_iter * K / (e2e_ms * 1e-3), "unit": "exchange-steps/s",
               "h2d_bytes_per_step": int(B * cfg.img_feat_dim * 4 + B * 8), "d2h_bytes_per_step": capi.MMG_LOSS_COUNT * 4,
               "ms_per_step": e2e_ms / K, "note": "mmg_host_prefetch + mmg_train_step_staged: pinned host x/target -> device on a "
               "copy stream (double-buffered, overlaps the previous step), losses -> pinned host every step; timed host-side "
               "from the first enqueue to the final synchronize"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bytes_iter = algorithmic_bytes_per_iteration(cfgd, world)
    # DRAM traffic per iteration from the committed ncu capture of this same command (profiles/, L2-flushed protocol)
    traffic, traffic_src = None, None
    if name in ("C2", "C2A") and world == 1 and flush is not None:
        try:
            import glob
            pat = "*_dram_traffic.json" if name == "C2" else "*_c2a_desc_attn_dram.json"
            f = sorted(glob.glob(os.path.join(ROOT, "profiles", pat)))[-1]
            traffic = float(json.load(open(f))["per_iteration"]["total_bytes"])
            traffic_src = os.path.relpath(f, ROOT)
        except Exception:
            pass
    ms_per_step = tot_ms / K
    achieved = bytes_iter / (ms_per_step * 1e-3) / 1e9
    value = world * steps_per_iter * K / (tot_ms * 1e-3)
    out = {
        "metric": "exchange-steps/sec", "value": value, "unit": "exchange-steps/s", "n_gpus": world, "steps": K,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[name], "batch_per_gpu": B, "global_batch": B * world, "exchange_steps": steps_per_iter,
                   "sampling": "on-device Philox4x32-10 (parity tests inject the reference's float64 uniforms instead)",
                   "parallelism": "dp%d" % world,
                   "dp_reduction": {"peer": "in-kernel sums over NVLink peer memory (statistics in k_lossgrad, gradient in "
                                            "k_peer_allreduce_norm); no collective call", "nccl": "torch.distributed NCCL all-reduce "
                                            "(statistics + flat gradient)", "none": "single GPU"}[dp_mode],
                   "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA events)" if flush is not None
                   else "not flushed (working set ~25 MB stays L2 resident)"},
        "value_l2_warm": world * steps_per_iter * K / (warm_ms * 1e-3), "ms_per_step_l2_warm": warm_ms / K,
        "gpu_launches": int(launches), "launches_per_step": launches / float(K),
        "wall_s_timed_region": t_wall, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_iteration": bytes_iter,
                     "kernel": "whole iteration = one launch sequence of %d kernels (k_pre, k_exchange_fwd, k_baseline_fwd, k_stats, "
                               "k_lossgrad, k_exchange_bwd, %sk_wgrad, k_reduce_norm, k_update); achieved = algorithmic bytes per "
                               % ((10, "k_attn_reduce, ") if cfgd.get("desc_attn") else (9, "")) +
                               "iteration / CUDA-event time per iteration; the path is latency/dependency-bound, see DESIGN.md"},
    }
    if e2e is not None:
        out["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        total, steps, threads = time_oracle(cfgd, 30, 3)
        out["cpu_baseline"] = {"value": steps / total, "unit": "exchange-steps/s", "cores": threads, "kind": "port",
                               "sample": "30 training iterations of the same workload (oracle/game_oracle.py, torch CPU fp32, "
                                         "%d threads, os.cpu_count=%d), %.1f ms/iteration" % (threads, os.cpu_count(), 1e3 * total / 30)}
    out_stream.write(json.dumps(out) + "\n")
    out_stream.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
