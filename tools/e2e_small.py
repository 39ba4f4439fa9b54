"""score_units end to end on small chromosomes (C1 shape, 2,000 bins): host-bound regime."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from peakachu_b200 import _lib, synth, shard
from peakachu_b200.forest import FlatForest
flat = FlatForest.load("bench_data/c2_forest.npz")
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ch = synth.make_chromosome("chr1", nb, seed=1234, depth=300.0, band=330)
n = ch.n
rowptr = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); t.numpy()[...] = a; return t
p_rp, p_w = pinned(rowptr), pinned(ch.weights)
p_d = pinned((ch.bin2 - ch.bin1).astype(np.uint16).view(np.uint8)); p_c = pinned(ch.count.astype(np.uint16).view(np.uint8))
class PinnedMap:
    def nbins(self, key): return n
    def weights(self, key, name): return p_w.numpy()
    def upper_pixels_csr16(self, key): return p_rp.numpy(), p_d.numpy().view(np.uint16), p_c.numpy().view(np.uint16)
def run(k, depth, shared):
    units = [("chr%d" % (i + 1), 0, n) for i in range(k)]
    return shard.score_units(PinnedMap(), units, flat, correct="weight", lower=6, upper=300, res=10000, device=0, min_prob=0.5,
                             depth=depth, shared_score_stream=shared)
for shared in (True, False):
    for depth in (1, 3, 6):
        run(2 * depth, depth, shared); torch.cuda.synchronize()
        t0 = time.perf_counter(); run(200, depth, shared); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("shared_score_stream=%s depth %d: %.3f ms per chromosome" % (shared, depth, dt / 200 * 1e3))
