import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from peakachu_b200 import _lib, synth, shard
from peakachu_b200.forest import FlatForest
flat = FlatForest.load("bench_data/c2_forest.npz")
ch = synth.make_chromosome("chr1", 24900, seed=1234, depth=300.0, band=330)
n = ch.n
rowptr = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); t.numpy()[...] = a; return t
p_rp, p_b2, p_cnt, p_w = pinned(rowptr), pinned(ch.bin2), pinned(ch.count), pinned(ch.weights)
class PinnedMap:
    def nbins(self, key): return n
    def weights(self, key, name): return p_w.numpy()
    def upper_pixels_csr(self, key): return p_rp.numpy(), p_b2.numpy(), p_cnt.numpy()
def run(k, depth):
    units = [("chr%d" % (i + 1), 0, n) for i in range(k)]
    return shard.score_units(PinnedMap(), units, flat, correct="weight", lower=6, upper=300, res=10000, device=0, min_prob=0.5, depth=depth)
for depth in (1, 2, 3, 4):
    run(4, depth); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(20, depth); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("depth %d: %.3f ms per chromosome" % (depth, dt / 20 * 1e3))
