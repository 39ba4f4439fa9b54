import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from peakachu_b200 import _lib, synth
from peakachu_b200.forest import FlatForest
from peakachu_b200.scoreUtils import Chromosome, DeviceForest
L = _lib.lib()
flat = FlatForest.load("bench_data/c2_forest.npz")
forest = DeviceForest.of(flat, 0)
ch = synth.make_chromosome("chr1", 24900, seed=1234, depth=300.0, band=330)
n = ch.n
rowptr = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); t.numpy()[...] = a; return t
p_rp, p_b2, p_cnt, p_w = pinned(rowptr), pinned(ch.bin2), pinned(ch.count), pinned(ch.weights)
def step(timing=None):
    t = [time.perf_counter()]
    X = Chromosome.from_csr(p_rp.numpy(), p_b2.numpy(), p_cnt.numpy(), p_w.numpy(), n, forest, lower=6, upper=300, cname="chr1", res=10000, width=5)
    t.append(time.perf_counter())
    _lib.check(L.pk_chrom_score(X._h, forest.handle, 0.5)); t.append(time.perf_counter())
    nrec = C.c_int64(); _lib.check(L.pk_chrom_result_count(X._h, C.byref(nrec), None, None)); t.append(time.perf_counter())
    m = nrec.value
    x, y = np.empty(m, np.int32), np.empty(m, np.int32); p, v = np.empty(m), np.empty(m)
    _lib.check(L.pk_chrom_fetch_results(X._h, _lib.ptr(x), _lib.ptr(y), _lib.ptr(p), _lib.ptr(v), None, m, 0)); t.append(time.perf_counter())
    X.close(); t.append(time.perf_counter())
    return np.diff(t) * 1e3
for _ in range(3): step()
acc = np.mean([step() for _ in range(10)], axis=0)
print("ms: from_csr(create+upload+fit+find launches) %.3f | score launch %.3f | result_count (sync) %.3f | fetch+sort %.3f | close %.3f | total %.3f" % (*acc, acc.sum()))
