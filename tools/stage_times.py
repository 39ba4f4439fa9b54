#!/usr/bin/env python
"""Per-stage times (CUDA events, pk_chrom_stage_ms) of one chromosome of a bench workload for every upload
encoding: which band-build kernel costs what. Median of --reps runs, one chromosome at a time."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    from peakachu_b200 import rowpack
    from peakachu_b200.forest import FlatForest
    from peakachu_b200.scoreUtils import Chromosome, DeviceForest
    wl = bench.WORKLOADS[args.workload]
    flat = FlatForest.load(os.path.join(ROOT, "bench_data", wl["forest"] + "_forest.npz"))
    forest = DeviceForest.of(flat, 0)
    ch = bench.make_map(wl, seed=1234)
    n = ch.n
    rp = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)
    kw = dict(lower=wl["lower"], upper=wl["upper"], cname="chr1", res=wl["res"], width=wl["w"])
    nd = (wl["upper"] + 2 * wl["w"] + 1 + 31) // 32 * 32
    blob = rowpack.pack_rows(rp, ch.bin2, ch.count, n, nd)
    makers = {
        "rows": lambda: Chromosome.from_rows(blob, ch.weights, n, forest, **kw),
        "csr16": lambda: Chromosome.from_csr16(rp, (ch.bin2 - ch.bin1).astype(np.uint16), ch.count.astype(np.uint16),
                                               ch.weights, n, forest, **kw),
        "csr32": lambda: Chromosome.from_csr(rp, ch.bin2, ch.count, ch.weights, n, forest, **kw),
    }
    for name, mk in makers.items():
        rows = []
        for _ in range(args.reps):
            X = mk()
            X.score_records(0.5)
            rows.append(X.stage_ms())
            X.close()
        med = {k: float(np.median([r[k] for r in rows])) for k in rows[0]}
        print(name, {k: round(1e3 * v, 1) for k, v in med.items()}, "us; sum %.1f us" % (1e3 * sum(med.values())))


if __name__ == "__main__":
    main()
