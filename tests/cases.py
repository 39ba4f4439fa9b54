"""Rebuild the inputs of a golden case (tests/golden/make_golden.py CASES) without
the reference: stored pixel arrays when the fixture carries them, otherwise the
seeded generator, guarded by the recorded checksum."""
import json
import os

import numpy as np

from peakachu_b200 import coolio, synth
from peakachu_b200.forest import FlatForest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case:
    def __init__(self, name):
        self.name = name
        with open(os.path.join(GOLDEN, name + ".json")) as fh:
            self.meta = json.load(fh)
        self.cfg = self.meta["case"]
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        pre = self.cfg["forest"].get("pretrained")          # a forest kept elsewhere in the repository (bench_data/)
        stem = os.path.join(os.path.dirname(os.path.dirname(GOLDEN)), pre) if pre else os.path.join(GOLDEN, name)
        self.pkl = stem + ".pkl"
        self.forest = FlatForest.load(stem + "_forest.npz")
        bed = os.path.join(GOLDEN, name + ".bedpe")          # compact cases keep a checksum and the records instead
        self.bedpe = open(bed).read() if os.path.exists(bed) else None
        self._chroms = None

    @property
    def chroms(self):
        if self._chroms is None:
            out = []
            for spec in self.cfg["chroms"]:
                nm = spec["name"]
                if spec.get("builder") == "batchrule":
                    ch = make_batchrule_chromosome(nm, spec["n"])
                elif self.cfg["store_inputs"]:
                    ch = synth.SynthChrom(name=nm, n=spec["n"], bin1=self.z[nm + "/bin1"],
                                          bin2=self.z[nm + "/bin2"], count=self.z[nm + "/count"],
                                          weights=self.z[nm + "/weights"], loops=np.zeros((0, 2), np.int64))
                else:
                    kw = {k: v for k, v in spec.items() if k not in ("name", "n")}
                    ch = synth.make_chromosome(nm, spec["n"], **kw)
                assert ch.checksum() == self.meta["checksums"][nm], \
                    "synthetic input for %s/%s does not match the fixture checksum" % (self.name, nm)
                out.append(ch)
            self._chroms = out
        return self._chroms

    def write_cool(self, tmpdir):
        path = os.path.join(str(tmpdir), self.name + ".pkcool")
        coolio.PKCool.write(path, self.chroms, self.cfg["res"])
        return path

    def model(self):
        import joblib
        return joblib.load(self.pkl)


    def pool(self, thr):
        """Output of the reference's `peakachu pool -t thr` on the golden bedpe (make_golden.py tap (vi))."""
        return open(os.path.join(GOLDEN, "%s.pool_t%s.bedpe" % (self.name, thr))).read()

    def chroms_arg(self):
        """--chroms of the score_genome run that made the fixture."""
        return self.cfg.get("chroms_arg", ["#", "X"])


# cases with every tap stored (tests parametrised over all taps)
ALL_CASES = ["tiny", "tiny_raw", "w7", "lowdepth", "c1", "genome", "c5", "gnames", "batchrule"]
# BASELINE configs[1] at full size: checksums of the large taps only
BIG_CASES = ["c2"]
FULL_TAP_CASES = ["tiny", "tiny_raw", "w7", "lowdepth"]


def make_batchrule_chromosome(name="chr7", n=14000):
    """A map for the reference's 100,000-candidate batch rule (scoreUtils.py:104-108): a batch in
    which at most one window survives the filters of utils.py:225-232 is dropped whole.

    A sparse lattice of count-2 pixels (x = 0 mod 6, d = 0 mod 2) makes ~345,000 Poisson candidates
    (four batches) whose windows hold at most six non-zero cells, so the 10 % filter rejects all of
    them. "Blobs" -- an 11 x 11 patch of count 1 around a count-25 centre, the lattice cleared
    around it -- add exactly one surviving window each (the count-1 cells are no candidates):
      batch 0: two blobs, 11,000 rows apart (kept; with three row tiles each tile sees only one)
      batch 1: one blob   (dropped by the rule)
      batch 2: none       (dropped)
      batch 3: three blobs (kept)
    Deterministic, no random numbers."""
    from peakachu_b200 import synth
    band = 320
    cnt = np.zeros((band, n), dtype=np.int32)                  # cnt[d, x] = pixel (x, x + d)
    xs = np.arange(0, n, 6)
    for d in range(0, band, 2):
        cnt[d, xs[xs + d < n]] = 2
    blobs = [(1000, 40), (12000, 40), (5000, 130), (2000, 284), (6000, 284), (9000, 284)]
    for xc, dc in blobs:
        yc = xc + dc
        # clear the lattice wherever a window could see the blob, then paint the blob
        for x in range(xc - 16, xc + 17):
            for y in range(yc - 16, yc + 17):
                if 0 <= y - x < band:
                    cnt[y - x, x] = 0
        for x in range(xc - 5, xc + 6):
            for y in range(yc - 5, yc + 6):
                cnt[y - x, x] = 1
        cnt[dc, xc] = 25
    dd, xx = np.nonzero(cnt)
    # one far pixel per bin (beyond the band) so that every bin owns a finite pixel (utils.py:151-156)
    far = np.arange(n - 1001)
    xx = np.concatenate([xx, far])
    dd = np.concatenate([dd, np.full(far.size, 1001)])
    cv = np.concatenate([cnt[dd[:-far.size], xx[:-far.size]], np.ones(far.size, dtype=np.int32)])
    order = np.lexsort((xx + dd, xx))
    b1 = xx[order].astype(np.int32)
    b2 = (xx + dd)[order].astype(np.int32)
    cc = cv[order].astype(np.int32)
    w = np.ones(n, dtype=np.float64)
    w[[300, 7001, 13500]] = np.nan                              # masked bins away from the blobs
    return synth.SynthChrom(name=name, n=n, bin1=b1, bin2=b2, count=cc, weights=w,
                            loops=np.array([[x, x + d] for x, d in blobs], dtype=np.int64))
