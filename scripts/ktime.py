#!/usr/bin/env python
"""Per-kernel device times of one training iteration, measured IN SITU (MMG_KTIME=1: CUDA events around every launch of the
library, no PDL) under bench.py's own two protocols: L2 flushed between iterations, and back to back (L2 warm).

    python scripts/ktime.py [--config C2] [--iters 60]

ncu's launch list replays every kernel serialised behind its own cache flush; this is the complementary view: the kernels in
their real order with whatever the previous kernel left in L2."""
import argparse
import ctypes as C
import os
import sys

os.environ["MMG_KTIME"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
import __graft_entry__ as ge


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--iters", type=int, default=60)
    args = ap.parse_args()
    ge.build()
    from multimodalgame_b200 import capi, engine as eng, synthetic as syn
    lib = capi.load()
    dev = torch.device("cuda", 0)
    fl = syn.GameFlags(**bench.CONFIGS[args.config])
    words = syn.desc_set(fl, seed=0)
    e = eng.GameEngine(syn.config_from_flags(fl, n_words=int(words["desc_set"].shape[0]) if words else 0), device=dev, lib=lib, seed=1)
    e.load_params(syn.init_params(fl, seed=0))
    if words:
        e.set_desc_set(**words)
    batches = [syn.batch(fl, seed=i) for i in range(4)]
    desc = batches[0][1].to(dev)
    xs = [b[0].to(dev) for b in batches]
    ts = [b[2].to(dev) for b in batches]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    buf = C.create_string_buffer(1 << 14)

    def report(tag):
        n = lib.dll.mmg_debug_kernel_times(buf, len(buf))
        rows = [l.split() for l in buf.value.decode().strip().splitlines()]
        tot = sum(float(r[2]) * int(r[1]) for r in rows) / max(1, args.iters)
        print("== %s: %d launches, %.1f us per iteration (sum of kernel times)" % (tag, n, tot))
        for r in rows:
            print("   %-34s n=%4d  %8.2f us" % (r[0], int(r[1]), float(r[2])))

    for i in range(20):
        e.train_step(xs[i % 4], desc, ts[i % 4])
    report("warm-up (discarded)")
    for i in range(args.iters):
        flush.zero_()
        e.train_step(xs[i % 4], desc, ts[i % 4])
    report("L2 flushed between iterations")
    for i in range(args.iters):
        e.train_step(xs[i % 4], desc, ts[i % 4])
    report("back to back (L2 warm)")


if __name__ == "__main__":
    main()
