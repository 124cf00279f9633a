#!/usr/bin/env python
"""A few training iterations of one bench.py configuration (for ncu / compute-sanitizer captures): python scripts/steps.py [--config C2] [--iters 6] [--flush]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
import __graft_entry__ as ge


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--flush", action="store_true")
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush else None
    for i in range(args.iters):
        if flush is not None:
            flush.zero_()
        e.train_step(xs[i % 4], desc, ts[i % 4])
    torch.cuda.synchronize()
    print("losses", {k: round(float(v), 5) for k, v in e.losses().items()})


if __name__ == "__main__":
    main()
