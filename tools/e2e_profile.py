"""cProfile of shard.score_units on small chromosomes (host-bound regime)."""
import sys, os, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from peakachu_b200 import _lib, synth, shard
from peakachu_b200.forest import FlatForest
flat = FlatForest.load("bench_data/c2_forest.npz")
ch = synth.make_chromosome("chr1", 2000, seed=1234, depth=300.0, band=330)
n = ch.n
rowptr = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)
d16, c16 = (ch.bin2 - ch.bin1).astype(np.uint16), ch.count.astype(np.uint16)
class Map:
    def nbins(self, key): return n
    def weights(self, key, name): return ch.weights
    def upper_pixels_csr16(self, key): return rowptr, d16, c16
def run(k):
    units = [("chr%d" % (i + 1), 0, n) for i in range(k)]
    return shard.score_units(Map(), units, flat, correct="weight", lower=6, upper=300, res=10000, device=0, min_prob=0.5)
run(12)
pr = cProfile.Profile(); pr.enable(); run(300); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
