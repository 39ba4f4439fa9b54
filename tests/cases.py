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
        self.pkl = os.path.join(GOLDEN, name + ".pkl")
        self.forest = FlatForest.load(os.path.join(GOLDEN, name + "_forest.npz"))
        self.bedpe = open(os.path.join(GOLDEN, name + ".bedpe")).read()
        self._chroms = None

    @property
    def chroms(self):
        if self._chroms is None:
            out = []
            for spec in self.cfg["chroms"]:
                nm = spec["name"]
                if self.cfg["store_inputs"]:
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


ALL_CASES = ["tiny", "tiny_raw", "w7", "lowdepth", "c1", "genome", "c5"]
FULL_TAP_CASES = ["tiny", "tiny_raw", "w7", "lowdepth"]
