#!/usr/bin/env python
"""What one rank of an N-GPU score_genome pass costs when it has the box to itself: the plan of `--world`
ranks is computed, the units of `--rank` are scored on this GPU alone (no PCIe / host-memory contention
from the other ranks), per-pass wall time and the share of the genome are printed. A lower bound for the
N-GPU pass time, measurable on one GPU."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", default="c3")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--ranks", default="0", help="comma list, or 'all'")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--encodings", default="rows,csr16,csr32")
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--largest-first", type=int, default=0)
    args = ap.parse_args()
    import torch
    from peakachu_b200 import shard, synth
    from peakachu_b200.forest import FlatForest
    wl = bench.GENOMES[args.genome]
    flat = FlatForest.load(os.path.join(ROOT, "bench_data", wl["forest"] + "_forest.npz"))
    sizes = synth.hg19_bins(wl["res"])
    queue = list(sizes)
    plan = shard.plan(sizes, args.world, wl["lower"], wl["upper"], wl["w"])
    px = sum(bench.band_pixels(n, wl["lower"], wl["upper"], wl["w"]) for n in sizes.values())
    ranks = range(args.world) if args.ranks == "all" else [int(r) for r in args.ranks.split(",")]
    encs = tuple(args.encodings.split(","))
    for r in ranks:
        mine = sorted(plan[r], key=lambda u: sizes[u[0]] * (u[2] - u[1]), reverse=bool(args.largest_first))
        pm = bench.PinnedMap(nd_enc=(wl["upper"] + 2 * wl["w"] + 1 + 31) // 32 * 32)
        for k in sorted({k for k, _, _ in mine}, key=queue.index):
            pm.add(k, synth.make_chromosome(k, sizes[k], seed=5000 + queue.index(k), depth=wl["depth"], band=wl["band"]),
                   formats=encs)
        share = sum(bench.band_pixels(sizes[k], wl["lower"], wl["upper"], wl["w"]) * (b - a) / sizes[k] for k, a, b in mine)
        for enc in encs:
            def one():
                return shard.score_units(pm, mine, flat, correct="weight", lower=wl["lower"], upper=wl["upper"],
                                         res=wl["res"], device=0, min_prob=0.5, copy=False, encoding=enc, depth=args.depth)
            for _ in range(3):
                one()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                one()
            torch.cuda.synchronize()
            ms = 1e3 * (time.perf_counter() - t0) / args.steps
            print("world %d rank %d: %d units %s, %.1f%% of the genome, %s: %.3f ms per pass (%.2e px/s on this share), h2d %.1f MB"
                  % (args.world, r, len(mine), [k for k, _, _ in mine], 100 * share / px, enc, ms, share / ms * 1e3,
                     sum(pm.h2d_bytes(k, enc) for k, _, _ in mine) / 1e6))


if __name__ == "__main__":
    main()
