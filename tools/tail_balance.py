#!/usr/bin/env python
"""A/B of pk_set_tuning("balance_tail"): the fused kernel's last round of candidates shared out evenly among the
CTAs (1) or first come first served (0). Fused-kernel time (stage 'features', median of --reps) on chromosomes of
several sizes; the records of both settings must be identical."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", default="2000,4813,9000,24900")
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    from peakachu_b200 import _lib, synth
    from peakachu_b200.forest import FlatForest
    from peakachu_b200.scoreUtils import Chromosome, DeviceForest
    L = _lib.lib()
    wl = bench.WORKLOADS["c2"]
    flat = FlatForest.load(os.path.join(ROOT, "bench_data", wl["forest"] + "_forest.npz"))
    forest = DeviceForest.of(flat, 0)
    for n in [int(x) for x in args.bins.split(",")]:
        ch = synth.make_chromosome("chr1", n, seed=1234, depth=wl["depth"], band=wl["band"])
        rp = np.searchsorted(ch.bin1, np.arange(ch.n + 1)).astype(np.int64)
        X = Chromosome.from_csr(rp, ch.bin2, ch.count, ch.weights, ch.n, forest, lower=wl["lower"], upper=wl["upper"],
                                cname="chr1", res=wl["res"], width=wl["w"])
        out = {}
        for knob in (0, 1, 0, 1):
            _lib.check(L.pk_set_tuning(b"balance_tail", knob))
            ts = []
            for _ in range(args.reps):
                rec = X.score_records(0.5)
                ts.append(X.stage_ms()["features"])
            out.setdefault(knob, []).append(1e3 * float(np.median(ts)))
            if "ref" not in out:
                out["ref"] = rec
            assert all(np.array_equal(a, b) for a, b in zip(out["ref"], rec)), "records differ"
        print("%6d bins, %7d candidates, %6d records: first come %s us, shared out %s us" % (
            n, X.n_candidates, out["ref"][0].size, ["%.1f" % t for t in out[0]], ["%.1f" % t for t in out[1]]))
        X.close()
    _lib.check(L.pk_set_tuning(b"balance_tail", 1))


if __name__ == "__main__":
    main()
