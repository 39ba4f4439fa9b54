#!/usr/bin/env python
"""Time the fused scoring kernel (stage 'features' of pk_chrom_stage_ms) for tuning variants of
pk_set_tuning("fused", v) on the bench chromosome; checks that every variant emits the same records."""
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
    ap.add_argument("--variants", default="-1,2,3,4")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--tma", type=int, default=2, help="pk_set_tuning('tma'): windows fetched as TMA boxes (default variant only)")
    args = ap.parse_args()
    from peakachu_b200 import _lib
    from peakachu_b200.forest import FlatForest
    from peakachu_b200.scoreUtils import Chromosome, DeviceForest
    L = _lib.lib()
    _lib.check(L.pk_set_tuning(b"tma", args.tma))
    wl = bench.WORKLOADS[args.workload]
    flat = FlatForest.load(os.path.join(ROOT, "bench_data", wl["forest"] + "_forest.npz"))
    forest = DeviceForest.of(flat, 0)
    ch = bench.make_map(wl, seed=1234)
    rp = np.searchsorted(ch.bin1, np.arange(ch.n + 1)).astype(np.int64)
    X = Chromosome.from_csr(rp, ch.bin2, ch.count, ch.weights, ch.n, forest, lower=wl["lower"], upper=wl["upper"],
                            cname="chr1", res=wl["res"], width=wl["w"])
    ref = None
    for v in [int(x) for x in args.variants.split(",")]:
        _lib.check(L.pk_set_tuning(b"fused", v))
        ts = []
        for _ in range(args.reps):
            rec = X.score_records(0.5)
            st = X.stage_ms()
            ts.append(st["features"])
        if ref is None:
            ref = rec
        same = all(np.array_equal(a, b) for a, b in zip(ref, rec))
        print("tma=%d fused=%d: %.1f us (min %.1f), records %d, identical %s" % (args.tma, v, 1e3 * float(np.median(ts)), 1e3 * min(ts), rec[0].size, same), "band_build %.1f us" % (1e3 * st["band_build"]))
    _lib.check(L.pk_set_tuning(b"fused", -1))
    X.close()


if __name__ == "__main__":
    main()
