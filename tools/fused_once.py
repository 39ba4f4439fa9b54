"""Score one C2- or C4-shaped chromosome a few times (used with -DPK_FUSED_CLOCK builds,
whose fused kernel prints per-phase cycle counts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peakachu_b200 import synth
from peakachu_b200.forest import FlatForest
from peakachu_b200.scoreUtils import Chromosome

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
from peakachu_b200 import _lib
if len(sys.argv) > 2:
    _lib.check(_lib.lib().pk_set_tuning(b"fused", int(sys.argv[2])))
if len(sys.argv) > 3:
    _lib.check(_lib.lib().pk_set_tuning(b"tma", int(sys.argv[3])))
if which == "c2":
    n, w, lower, upper, forest = 24900, 5, 6, 300, "bench_data/c2_forest.npz"
else:
    n, w, lower, upper, forest = 49850, 7, 6, 600, "bench_data/c4_forest.npz"
ch = synth.make_chromosome("chr1", n, seed=1234, depth=300.0, band=upper + 30)
flat = FlatForest.load(forest)
X = Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, ch.weights, ch.n, flat, lower=lower, upper=upper,
                           cname="chr1", res=10000, width=w, sorted_pixels=True)
for _ in range(3):
    X.score_records(0.5)
print(which, {k: round(v, 4) for k, v in X.stage_ms().items()})
X.close()
