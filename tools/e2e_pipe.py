"""shard.score_units end to end (narrow pinned columns) at several pipeline depths."""
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
p_rp, p_w = pinned(rowptr), pinned(ch.weights)
p_d = pinned((ch.bin2 - ch.bin1).astype(np.uint16).view(np.uint8)); p_c = pinned(ch.count.astype(np.uint16).view(np.uint8))
class PinnedMap:
    def nbins(self, key): return n
    def weights(self, key, name): return p_w.numpy()
    def upper_pixels_csr16(self, key): return p_rp.numpy(), p_d.numpy().view(np.uint16), p_c.numpy().view(np.uint16)
def run(k, depth):
    units = [("chr%d" % (i + 1), 0, n) for i in range(k)]
    return shard.score_units(PinnedMap(), units, flat, correct="weight", lower=6, upper=300, res=10000, device=0, min_prob=0.5, depth=depth)
for reserve in (0, 2, 4, 8, 12, 16):
    _lib.check(_lib.lib().pk_set_tuning(b"reserve_sms", reserve))
    for depth in (4, 6, 8):
        run(2 * depth, depth); torch.cuda.synchronize()
        t0 = time.perf_counter(); run(100, depth); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("reserve %2d depth %d: %.3f ms per chromosome" % (reserve, depth, dt / 100 * 1e3))
