#!/usr/bin/env python
"""Train the forests used by bench.py through the REFERENCE's own training-feature
code (trainUtils.buildmatrix, imported from /root/reference), on a synthetic
chromosome drawn from the same generator as the benchmark map but a different seed.
Run in the build container only; the fitted forests are committed here because
/root/reference does not exist on the GPU box.

  c2:  w=5 (121 features), 100 trees, max_depth 20   -- BASELINE configs[1], [2]
  c4:  w=7 (225 features), 200 trees, max_depth 25   -- BASELINE configs[3]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests", "golden"))
import make_golden as mg  # noqa: E402  (sets up the cooler stand-in and imports the reference)
import joblib  # noqa: E402
from peakachu_b200.forest import flatten_forest  # noqa: E402

MODELS = {
    "c2": dict(train=dict(name="chrT", n=24900, seed=777, depth=300.0, band=330, n_loops=6000),
               res=10000, w=5, weight="weight", upper=300,
               forest=dict(n_estimators=100, max_depth=20, seed=0)),
    "c4": dict(train=dict(name="chrT", n=30000, seed=778, depth=300.0, band=640, n_loops=6000, loop_max=550),
               res=5000, w=7, weight="weight", upper=600,
               forest=dict(n_estimators=200, max_depth=25, seed=0)),
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(MODELS)):
        model, shape = mg.train_forest(MODELS[name])
        ff = flatten_forest(model)
        print(name, "trained on", shape, "nodes", ff.n_nodes, "nodes/tree", ff.n_nodes / ff.n_trees)
        joblib.dump(model, os.path.join(HERE, name + ".pkl"), compress=("xz", 3))
        ff.save(os.path.join(HERE, name + "_forest.npz"))
