"""Flatten a fitted scikit-learn ``RandomForestClassifier`` into plain arrays.

The reference scores windows with ``model.predict_proba(fea)[:, 1]``
(``scoreUtils.py:109``) on a joblib-loaded forest (``score_chromosome.py:14``).
The CUDA path needs the same trees as node tables. What sklearn does at predict
time (``ensemble/_forest.py`` ``predict_proba`` + ``tree/_tree.pyx``
``_apply_dense``):

* X is cast to float32;
* per tree, from node 0: ``isnan(x) ? missing_go_to_left : (x_f32 <= threshold_f64)``
  picks left/right until ``left_child == -1``;
* ``acc_f64 += tree_.value[leaf, 0, 1]`` in ``estimators_`` order, then
  ``acc /= n_estimators``.

``FlatForest`` carries those tables, nothing else. Thresholds stay float64 here;
the C library narrows them (round toward -inf to float32, which preserves
``x_f32 <= t`` exactly) when it packs nodes for the device.
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class FlatForest:
    n_trees: int
    n_features: int
    node_offset: np.ndarray   # int64[n_trees + 1] first node of each tree
    feature: np.ndarray       # int32[n_nodes], < 0 for leaves
    threshold: np.ndarray     # float64[n_nodes]
    left: np.ndarray          # int32[n_nodes] tree-local child index, -1 for leaves
    right: np.ndarray         # int32[n_nodes]
    missing_left: np.ndarray  # uint8[n_nodes]
    leaf_p1: np.ndarray       # float64[n_nodes] P(class index 1) at that node

    @property
    def width(self) -> int:
        """Window half-width the forest was trained with
        (``score_chromosome.py:23``)."""
        return int((np.sqrt(self.n_features) - 1) / 2)

    @property
    def n_nodes(self) -> int:
        return int(self.node_offset[-1])

    def save(self, path: str) -> None:
        np.savez_compressed(path, **{f.name: getattr(self, f.name)
                                     for f in dataclasses.fields(self)})

    @staticmethod
    def load(path: str) -> "FlatForest":
        z = np.load(path)
        return FlatForest(n_trees=int(z["n_trees"]), n_features=int(z["n_features"]),
                          **{k: z[k] for k in ("node_offset", "feature", "threshold", "left",
                                               "right", "missing_left", "leaf_p1")})


def flatten_forest(model) -> FlatForest:
    """sklearn RandomForestClassifier (binary, single output) -> FlatForest."""
    ests = getattr(model, "estimators_", None)
    if ests is None:
        raise TypeError("model has no estimators_: not a fitted forest")
    if getattr(model, "n_outputs_", 1) != 1 or int(np.atleast_1d(model.n_classes_)[0]) != 2:
        raise ValueError("scoring path needs a single-output two-class forest")
    feats, thrs, lefts, rights, miss, p1s, offs = [], [], [], [], [], [], [0]
    for est in ests:
        t = est.tree_
        val = np.asarray(t.value, dtype=np.float64)[:, 0, :]
        # what DecisionTreeClassifier.predict_proba returns for a sample landing
        # on this node: tree_.value holds class fractions since sklearn 1.3; older
        # pickles hold weighted counts and are normalised at predict time.
        s = val.sum(axis=1)
        if not np.all(np.abs(s - 1.0) < 1e-9):
            val = val / np.where(s == 0.0, 1.0, s)[:, None]
        feats.append(np.asarray(t.feature, dtype=np.int32))
        thrs.append(np.asarray(t.threshold, dtype=np.float64))
        lefts.append(np.asarray(t.children_left, dtype=np.int32))
        rights.append(np.asarray(t.children_right, dtype=np.int32))
        mgl = getattr(t, "missing_go_to_left", None)
        miss.append(np.zeros(t.node_count, np.uint8) if mgl is None
                    else np.asarray(mgl, dtype=np.uint8))
        p1s.append(np.ascontiguousarray(val[:, 1]))
        offs.append(offs[-1] + t.node_count)
    return FlatForest(
        n_trees=len(ests), n_features=int(model.n_features_in_),
        node_offset=np.asarray(offs, dtype=np.int64),
        feature=np.concatenate(feats), threshold=np.concatenate(thrs),
        left=np.concatenate(lefts), right=np.concatenate(rights),
        missing_left=np.concatenate(miss), leaf_p1=np.concatenate(p1s))


def load_model(path: str):
    """``joblib.load`` like the reference, or a ``.npz`` FlatForest dump.
    Returns (FlatForest, sklearn model or None)."""
    if path.endswith(".npz"):
        return FlatForest.load(path), None
    import joblib
    model = joblib.load(path)
    return flatten_forest(model), model
