"""Golden vectors for trainUtils.buildmatrix (trainUtils.py:12-44), made by running the
reference itself (read-only tree at /root/reference) on the inputs of three golden cases.
Run in the build container:  python tests/golden/make_buildmatrix_golden.py
Writes tests/golden/buildmatrix.npz: <case>/coords int32[n][2], <case>/fea float64[kept][F]."""
import os
import sys
import tempfile

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402

from peakachu_b200 import coolio  # noqa: E402

sys.modules["cooler"] = coolio          # the reference's four cooler calls, served by the stand-in
from peakachu import trainUtils, utils  # noqa: E402  (the reference)
from cases import Case  # noqa: E402

N_COORDS = 150


def main():
    out = {}
    for name in ("tiny", "tiny_raw", "w7"):
        cs = Case(name)
        ch, cfg = cs.chroms[0], cs.cfg
        lib = coolio.Cooler(cs.write_cool(tempfile.mkdtemp()))
        balance = False if cfg["weight"] == "raw" else cfg["weight"]
        M = utils.tocsr(lib.matrix(balance=balance, sparse=True).fetch(ch.name))
        rng = np.random.default_rng(11)
        d = ch.bin2 - ch.bin1
        ok = np.nonzero((d > 0) & (d <= cfg["upper"]))[0]
        pick = rng.choice(ok, size=N_COORDS, replace=False)
        coords = [(int(ch.bin1[i]), int(ch.bin2[i])) for i in pick]
        coords += [(3, 40), (ch.n - 4, ch.n - 2), (10, 12)]      # masked out by trainUtils.py:22
        fea = trainUtils.buildmatrix(M, coords, w=cfg["w"])
        out[name + "/coords"] = np.array(coords, np.int32)
        out[name + "/fea"] = np.array(fea)
        print(name, len(coords), "coords ->", len(fea), "windows")
    np.savez_compressed(os.path.join(HERE, "buildmatrix.npz"), **out)


if __name__ == "__main__":
    main()
