#!/usr/bin/env python
"""Golden output of the reference's `peakachu depth` (calculate_depth.py, UNMODIFIED, imported from
/root/reference) on the maps of the golden cases: the three lines it prints -- total intra-chromosomal
contacts, the human-equivalent depth and the suggested pre-trained model. BASELINE.json configs[4] picks the
forest from that suggestion; tests/test_gpu_parity.py::test_depth_selects_the_model_like_the_reference replays
the step on the GPU path. Run in the build container only:  python tests/golden/make_depth_golden.py"""
import argparse
import io
import json
import os
import sys
import tempfile
import types
from contextlib import redirect_stdout

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from peakachu_b200 import coolio  # noqa: E402

sys.modules["cooler"] = types.ModuleType("cooler")
sys.modules["cooler"].Cooler = coolio.Cooler          # `cooler` is absent from the image (see make_golden.py)

from peakachu import calculate_depth  # noqa: E402
from tests.cases import Case  # noqa: E402

out = {}
for name, min_dis in [("c5", 0), ("c5", 20000), ("genome", 0), ("gnames", 30000), ("tiny", 0), ("lowdepth", 50000)]:
    case = Case(name)
    path = case.write_cool(tempfile.mkdtemp())
    buf = io.StringIO()
    with redirect_stdout(buf):
        calculate_depth.main(argparse.Namespace(path=path, min_dis=min_dis))
    out.setdefault(name, {})[str(min_dis)] = buf.getvalue()
    print(name, min_dis, repr(buf.getvalue()[-160:]))
with open(os.path.join(HERE, "depth.json"), "w") as fh:
    json.dump(out, fh, indent=1)
