#!/usr/bin/env python
"""Where the wall time of `score_chromosome.main(args)` goes from a file path to the bedpe on disk
(cProfile after a warm-up run; the c2 chromosome as .cool and .pkcool)."""
import argparse
import contextlib
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from peakachu_b200 import coolio, score_chromosome
    from tests import h5write
    wl = bench.WORKLOADS["c2"]
    ch = bench.make_map(wl, 1234)
    tmp = tempfile.mkdtemp(prefix="pk_file_prof_")
    paths = {"cool": os.path.join(tmp, "wl.cool"), "pkcool": os.path.join(tmp, "wl.pkcool")}
    h5write.write_cool(paths["cool"], [ch], wl["res"])
    coolio.PKCool.write(paths["pkcool"], [ch], wl["res"])
    for kind, path in paths.items():
        ns = argparse.Namespace(path=path, model=os.path.join(ROOT, "bench_data", wl["forest"] + ".pkl"), output=os.path.join(tmp, kind + ".bedpe"),
                                resolution=wl["res"], lower=wl["lower"], upper=wl["upper"], minimum_prob=0.5,
                                clr_weight_name="weight", chrom=ch.name, device=0)
        for _ in range(2):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                score_chromosome.main(ns)
            print(kind, "run %.1f ms" % (1e3 * (time.perf_counter() - t0)))
        pr = cProfile.Profile()
        pr.enable()
        with contextlib.redirect_stdout(io.StringIO()):
            score_chromosome.main(ns)
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(28)
        print(kind, s.getvalue()[:6000])


if __name__ == "__main__":
    main()
