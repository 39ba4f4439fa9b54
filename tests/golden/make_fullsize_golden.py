#!/usr/bin/env python
"""Full-size golden checksums for BASELINE.json configs[2] and configs[3], from the UNMODIFIED reference
(imported from /root/reference; `cooler` served by the stand-in, see make_golden.py):

  c3       score_genome.main on every chromosome of the hg19-shaped 10 kb genome bench.py scores
           (synth.make_chromosome(name, hg19_bins(10000)[name], seed=5000 + index, depth=300, band=330),
           bench_data/c2.pkl, -l 6 -u 300, --minimum-prob 0.5): sha256 and row count of each chromosome's bedpe
  c4_chr1  score_chromosome.main on bench.py's c4 chromosome (49,850 bins at 5 kb, seed 1234, bench_data/c4.pkl,
           w = 7, -u 600): sha256 and row count of the bedpe

Chromosomes go through the reference one file at a time (a 23-chromosome container would not change its
arithmetic: score_genome.main scores them one by one, score_genome.py:46-84). About 15 minutes of one core.
    python tests/golden/make_fullsize_golden.py [c3] [c4_chr1]        ->  tests/golden/fullsize.json"""
import argparse
import hashlib
import io
import json
import os
import sys
import tempfile
import time
import types
from contextlib import redirect_stdout

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

from peakachu_b200 import coolio, synth  # noqa: E402

sys.modules["cooler"] = types.ModuleType("cooler")
sys.modules["cooler"].Cooler = coolio.Cooler

from peakachu import score_chromosome, score_genome  # noqa: E402

import bench  # noqa: E402

np.seterr(divide="ignore", invalid="ignore")
OUT = os.path.join(HERE, "fullsize.json")
want = sys.argv[1:] or ["c3", "c4_chr1"]
res = json.load(open(OUT)) if os.path.exists(OUT) else {}


def run(main, ns):
    with redirect_stdout(io.StringIO()):
        main(ns)
    txt = open(ns.output).read()
    return dict(sha256=hashlib.sha256(txt.encode()).hexdigest(), rows=txt.count("\n"))


if "c3" in want:
    wl = bench.GENOMES["c3"]
    sizes = synth.hg19_bins(wl["res"])
    res["c3"] = dict(workload=wl["desc"], chroms={})
    for idx, (name, n) in enumerate(sizes.items()):
        t0 = time.time()
        ch = synth.make_chromosome(name, n, seed=5000 + idx, depth=wl["depth"], band=wl["band"])
        tmp = tempfile.mkdtemp()
        path = os.path.join(tmp, name + ".pkcool")
        coolio.PKCool.write(path, [ch], wl["res"], rows_nd=0)
        ns = argparse.Namespace(path=path, model=os.path.join(ROOT, "bench_data", wl["forest"] + ".pkl"),
                                output=os.path.join(tmp, "o.bedpe"), resolution=wl["res"], lower=wl["lower"], upper=wl["upper"],
                                minimum_prob=0.5, clr_weight_name="weight", chroms=[])
        r = run(score_genome.main, ns)
        r.update(n_bins=n, seed=5000 + idx, input_checksum=ch.checksum())
        res["c3"]["chroms"][name] = r
        print("c3", name, n, r["rows"], "%.0f s" % (time.time() - t0), flush=True)
        json.dump(res, open(OUT, "w"), indent=1)
        os.remove(path)

if "c4_chr1" in want:
    wl = bench.WORKLOADS["c4"]
    t0 = time.time()
    ch = bench.make_map(wl, seed=1234)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "c4.pkcool")
    coolio.PKCool.write(path, [ch], wl["res"], rows_nd=0)
    ns = argparse.Namespace(path=path, model=os.path.join(ROOT, "bench_data", wl["forest"] + ".pkl"),
                            output=os.path.join(tmp, "o.bedpe"), resolution=wl["res"], lower=wl["lower"], upper=wl["upper"],
                            minimum_prob=0.5, clr_weight_name="weight", chrom=ch.name)
    r = run(score_chromosome.main, ns)
    r.update(n_bins=ch.n, seed=1234, input_checksum=ch.checksum(), workload=wl["desc"])
    res["c4_chr1"] = r
    print("c4_chr1", r["rows"], "%.0f s" % (time.time() - t0), flush=True)
    json.dump(res, open(OUT, "w"), indent=1)
