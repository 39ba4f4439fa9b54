"""Where the end-to-end time of shard.score_units goes on the host: per chromosome the
submit calls, the wait for the device, and the result fetch (narrow columns, depth 3)."""
import sys, os, time, ctypes as C
from collections import deque
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from peakachu_b200 import _lib, synth, shard
from peakachu_b200.forest import FlatForest
from peakachu_b200.scoreUtils import Chromosome, DeviceForest
flat = FlatForest.load("bench_data/c2_forest.npz")
ch = synth.make_chromosome("chr1", 24900, seed=1234, depth=300.0, band=330)
n = ch.n
rowptr = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); t.numpy()[...] = a; return t
p_rp, p_w = pinned(rowptr), pinned(ch.weights)
p_d = pinned((ch.bin2 - ch.bin1).astype(np.uint16).view(np.uint8)); p_c = pinned(ch.count.astype(np.uint16).view(np.uint8))
d16, c16 = p_d.numpy().view(np.uint16), p_c.numpy().view(np.uint16)
L = _lib.lib()
forest = DeviceForest.of(flat, 0)
streams = []
for _ in range(3):
    st = C.c_void_p(); _lib.check(L.pk_stream_create(0, C.byref(st))); streams.append(st)
T = dict(submit=0.0, wait=0.0, fetch=0.0, close=0.0)
STAGES = None
def submit(i):
    t0 = time.perf_counter()
    X = Chromosome.from_csr16(p_rp.numpy(), d16, c16, p_w.numpy(), n, forest, lower=6, upper=300, cname="chr1", res=10000,
                              width=5, device=0, stream=streams[i % 3].value, first_tile=(0, n))
    _lib.check(L.pk_chrom_score(X._h, forest.handle, 0.5))
    T["submit"] += time.perf_counter() - t0
    return X
def finish(X):
    t0 = time.perf_counter()
    nrec, nc = C.c_int64(), C.c_int64()
    _lib.check(L.pk_chrom_result_count(X._h, C.byref(nrec), C.byref(nc), None))
    t1 = time.perf_counter()
    shard._fetch_tile(X, 0, n, n)
    t2 = time.perf_counter()
    if STAGES is not None:
        for k, v in X.stage_ms().items():
            STAGES[k] = STAGES.get(k, 0.0) + v
    X.close()
    t3 = time.perf_counter()
    T["wait"] += t1 - t0; T["fetch"] += t2 - t1; T["close"] += t3 - t2
def run(k):
    q = deque()
    for i in range(k):
        if len(q) == 3:
            finish(q.popleft())
        q.append(submit(i))
    while q:
        finish(q.popleft())
run(6); torch.cuda.synchronize()
for k in T: T[k] = 0.0
K = 100
t0 = time.perf_counter(); run(K); torch.cuda.synchronize(); dt = time.perf_counter() - t0
STAGES = {}
run(30); torch.cuda.synchronize()
print("stage ms inside the pipeline:", {k: round(v / 30, 4) for k, v in STAGES.items()}, "sum %.3f" % (sum(STAGES.values()) / 30))
print("total %.3f ms per chromosome; host: " % (dt / K * 1e3) + ", ".join("%s %.3f" % (k, v / K * 1e3) for k, v in T.items()))
