#!/usr/bin/env python
"""Per-CTA timeline of one training iteration from a -DMMG_TRACE debug build (scripts/dbg/libmmg_trace.so, built here on
first use): python scripts/trace.py [--config C2] [--pdl 0|1].  Prints, per kernel, the span from the first CTA start to the
last CTA end and the distribution of the per-CTA stamp slots (all relative to the first stamp of the iteration)."""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import __graft_entry__ as ge

NAMES = ["k_pre", "k_exchange_fwd", "k_baseline_fwd", "k_exchange_bwd", "k_wgrad", "k_update", "k_peer_reduce_scatter"]


def build_trace_lib():
    out = os.path.join(ROOT, "scripts", "dbg", "libmmg_trace.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.isfile(out) or os.path.getmtime(out) < ge._newest_source_mtime():
        flags = [f for f in ge.NVCC_FLAGS if not f.startswith("--use_fast_math")]
        subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + flags + ["-DMMG_TRACE", os.path.join(ge.CSRC, "mmg_api.cu"), "-o", out])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--build-only", action="store_true")
    ap.add_argument("--dump", type=int, default=-1, help="print every CTA of kernel index K (last iteration)")
    args = ap.parse_args()
    path = build_trace_lib()
    if args.build_only:
        return
    from multimodalgame_b200 import capi, engine as eng, synthetic as syn
    lib = capi.Library(path)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    fl = syn.GameFlags(**bench.CONFIGS[args.config])
    words = syn.desc_set(fl, seed=0)
    B = fl.batch_size
    e = eng.GameEngine(syn.config_from_flags(fl, batch_global=B * world, batch_offset=rank * B,
                                             n_words=int(words["desc_set"].shape[0]) if words else 0), device=dev, lib=lib, seed=1)
    e.load_params(syn.init_params(fl, seed=0))
    if world > 1:       # torchrun: the NVLink peer-memory data-parallel iteration, rank 0 prints its timeline
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        e.enable_peer_dp()
        e.train_step = lambda x, d, t: e.train_step_peer(x, d, t)
    if words:
        e.set_desc_set(**words)
    batches = [syn.batch(fl, seed=10 * rank + i) for i in range(4)]
    desc = batches[0][1].to(dev)
    xs = [b[0].to(dev) for b in batches]
    ts = [b[2].to(dev) for b in batches]
    buf = np.zeros(7 * 1024 * 8, dtype=np.uint64)
    for i in range(30):
        e.train_step(xs[i % 4], desc, ts[i % 4])
    lib.dll.mmg_debug_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
    for rep in range(3):
        e.train_step(xs[rep % 4], desc, ts[rep % 4])
        n = lib.dll.mmg_debug_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
        assert n == buf.size, "not a -DMMG_TRACE build"
        tr = buf.reshape(7, 1024, 8).astype(np.float64)
        t0 = tr[tr > 0].min()
        if rank != 0:
            continue
        print("== iteration %d (us relative to the first stamp)" % rep)
        for k in range(7):
            tk = tr[k]
            used = (tk > 0).any(1)
            if not used.any():
                continue
            v = np.where(tk > 0, (tk - t0) * 1e-3, np.nan)
            print("  %-16s CTAs %4d  span %7.2f .. %7.2f us" % (NAMES[k], int(used.sum()), np.nanmin(v), np.nanmax(v)))
            for slot in range(8):
                col = v[:, slot]
                m = ~np.isnan(col)
                if m.any():
                    c = col[m]
                    print("      slot %d: n=%4d  min %7.2f  median %7.2f  max %7.2f" % (slot, int(m.sum()), c.min(), np.median(c), c.max()))
            if rep == 2 and k == args.dump:
                for cta in range(1024):
                    if used[cta]:
                        print("      cta %4d: %s" % (cta, " ".join("%7.2f" % x if not np.isnan(x) else "      ." for x in v[cta])))


if __name__ == "__main__":
    main()
