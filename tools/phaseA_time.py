import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from peakachu_b200 import _lib, synth
from peakachu_b200.forest import FlatForest
from peakachu_b200.scoreUtils import Chromosome, DeviceForest
L = _lib.lib()
ch = synth.make_chromosome("chr1", 24900, seed=1234, depth=300.0, band=330)
for name in ("tests/golden/tiny_forest.npz", "bench_data/c2_forest.npz"):
    flat = FlatForest.load(name)
    X = Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, ch.weights, ch.n, flat, lower=6, upper=300, cname="chr1", res=10000, width=5, sorted_pixels=True)
    for _ in range(4):
        X.score_records(0.5)
    print(name, flat.n_trees, "trees", flat.n_nodes, "nodes ->", {k: round(v, 4) for k, v in X.stage_ms().items()})
    X.close()
