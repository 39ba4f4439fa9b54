#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (imported from /root/reference) on seeded synthetic inputs.

Run here (build container) only:  python tests/golden/make_golden.py [case ...]

The reference has no tests, golden vectors or fixtures of its own (SURVEY.md
section 4), so the oracle under ``oracle/`` is pinned differentially against the
outputs recorded here, at six taps of the scoring path:

  (i)   ``Chromosome.ridx/cidx``  -- the Poisson candidate set (scoreUtils.py:40-68)
  (ii)  ``Chromosome.exp_arr``    -- expected curve (utils.py:139-178)
  (iii) ``Chromosome.getwindow``  -- kept coordinates + features (scoreUtils.py:70-93)
  (iv)  ``estimator.apply`` leaves and ``predict_proba`` (scoreUtils.py:109)
  (v)   bedpe text written by ``score_chromosome.main`` / ``score_genome.main``
  (vi)  ``peakachu pool`` output on that bedpe (call_loops.py:3-26)

``cooler`` is absent from the image; the reference's ``import cooler`` is served by
the stand-in ``peakachu_b200.coolio`` (see its docstring). ``peakachu train`` is
broken on Python >= 3.11 (trainUtils.py:150), so forests are fitted on features
from the reference's own ``trainUtils.buildmatrix`` with a directly constructed
``RandomForestClassifier`` (the estimator class trainUtils.trainRF grid-searches).
"""
from __future__ import annotations

import argparse
import hashlib
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

import numpy as np  # noqa: E402
import joblib  # noqa: E402

from peakachu_b200 import coolio, synth  # noqa: E402
from peakachu_b200.forest import flatten_forest  # noqa: E402

sys.modules["cooler"] = types.ModuleType("cooler")
sys.modules["cooler"].Cooler = coolio.Cooler

from peakachu import scoreUtils, trainUtils, utils  # noqa: E402
from peakachu import score_chromosome, score_genome, call_loops  # noqa: E402
from sklearn.ensemble import RandomForestClassifier  # noqa: E402

np.seterr(divide="ignore", invalid="ignore")


# --------------------------------------------------------------------------
# case table: everything a test needs to rebuild the inputs
# --------------------------------------------------------------------------
CASES = {
    # name: dict(score chromosomes, training chromosome, flags, forest)
    "tiny": dict(
        chroms=[dict(name="chr1", n=400, seed=11, depth=300.0, band=110, n_loops=30, loop_max=70)],
        train=dict(name="chrT", n=2000, seed=12, depth=300.0, band=110, n_loops=500, loop_max=70),
        lower=6, upper=60, res=10000, w=5, weight="weight", min_prob=0.5,
        forest=dict(n_estimators=20, max_depth=8, seed=1), store_inputs=True, full_taps=True),
    "tiny_raw": dict(
        chroms=[dict(name="chr1", n=400, seed=11, depth=300.0, band=110, n_loops=30, loop_max=70)],
        train=dict(name="chrT", n=2000, seed=12, depth=300.0, band=110, n_loops=500, loop_max=70),
        lower=6, upper=60, res=10000, w=5, weight="raw", min_prob=0.5,
        forest=dict(n_estimators=20, max_depth=8, seed=1), store_inputs=True, full_taps=True),
    "w7": dict(
        chroms=[dict(name="chr3", n=500, seed=21, depth=300.0, band=130, n_loops=40, loop_max=80)],
        train=dict(name="chrT", n=2000, seed=22, depth=300.0, band=130, n_loops=500, loop_max=80),
        lower=6, upper=80, res=5000, w=7, weight="weight", min_prob=0.5,
        forest=dict(n_estimators=30, max_depth=10, seed=2), store_inputs=True, full_taps=True),
    "lowdepth": dict(
        chroms=[dict(name="chr1", n=600, seed=31, depth=8.0, band=110, n_loops=50, loop_max=70)],
        train=dict(name="chrT", n=3000, seed=32, depth=8.0, band=110, n_loops=800, loop_max=70),
        lower=6, upper=60, res=10000, w=5, weight="weight", min_prob=0.3,
        forest=dict(n_estimators=20, max_depth=8, seed=3), store_inputs=True, full_taps=True),
    # BASELINE.json configs[0]: 2,000-bin 10 kb map, defaults, 100-tree forest
    "c1": dict(
        chroms=[dict(name="chr1", n=2000, seed=0, depth=300.0, band=330)],
        train=dict(name="chrT", n=6000, seed=1, depth=300.0, band=330, n_loops=1500),
        lower=6, upper=300, res=10000, w=5, weight="weight", min_prob=0.5,
        forest=dict(n_estimators=100, max_depth=20, seed=0), store_inputs=False, full_taps=False),
    # score_genome: chromosome filter ('#' + X by default), chr-prefixing, file order
    "genome": dict(
        chroms=[dict(name="1", n=700, seed=41, depth=300.0, band=130, n_loops=40, loop_max=80),
                dict(name="2", n=450, seed=42, depth=300.0, band=130, n_loops=25, loop_max=80),
                dict(name="X", n=520, seed=43, depth=300.0, band=130, n_loops=30, loop_max=80),
                dict(name="M", n=300, seed=44, depth=300.0, band=130, n_loops=10, loop_max=80)],
        train=dict(name="chrT", n=2500, seed=45, depth=300.0, band=130, n_loops=600, loop_max=80),
        lower=6, upper=80, res=10000, w=5, weight="weight", min_prob=0.5,
        forest=dict(n_estimators=40, max_depth=12, seed=4), store_inputs=False, full_taps=False,
        genome=True),
    # BASELINE.json configs[4]: score_genome on a low-depth map with a forest fitted at the matching depth
    # (what the reference's `depth` step selects), --minimum-prob 0.6, then `pool` at 0.6 and at the default 0.9
    "c5": dict(
        chroms=[dict(name="1", n=800, seed=51, depth=8.0, band=110, n_loops=60, loop_max=70),
                dict(name="5", n=550, seed=52, depth=8.0, band=110, n_loops=40, loop_max=70),
                dict(name="X", n=480, seed=53, depth=8.0, band=110, n_loops=30, loop_max=70)],
        train=dict(name="chrT", n=3000, seed=54, depth=8.0, band=110, n_loops=800, loop_max=70),
        lower=6, upper=60, res=10000, w=5, weight="weight", min_prob=0.6,
        forest=dict(n_estimators=60, max_depth=14, seed=5), store_inputs=False, full_taps=False,
        genome=True),
    # score_genome's labels (score_genome.py:48-51 prepends 'chr' unless the name starts with it; it does
    # not strip characters the way score_chromosome.py:37-38 does) on contig names that tell the two rules
    # apart; an empty --chroms list scores every chromosome (score_genome.py:43)
    "gnames": dict(
        chroms=[dict(name="chr1", n=420, seed=61, depth=300.0, band=110, n_loops=25, loop_max=70),
                dict(name="hs37d5", n=380, seed=62, depth=300.0, band=110, n_loops=20, loop_max=70),
                dict(name="contig1", n=350, seed=63, depth=300.0, band=110, n_loops=20, loop_max=70),
                dict(name="chrrDNA", n=330, seed=64, depth=300.0, band=110, n_loops=20, loop_max=70)],
        train=dict(name="chrT", n=2000, seed=65, depth=300.0, band=110, n_loops=500, loop_max=70),
        lower=6, upper=60, res=10000, w=5, weight="weight", min_prob=0.5,
        forest=dict(n_estimators=20, max_depth=8, seed=6), store_inputs=False, full_taps=False,
        genome=True, chroms_arg=[]),
    # the 100,000-candidate batch rule (scoreUtils.py:104-108): ~340,000 candidates in four batches that keep
    # 2 / 1 / 0 / 3 windows (tests/cases.py make_batchrule_chromosome); the reference drops the second batch's
    # only window
    "batchrule": dict(
        chroms=[dict(name="chr7", n=14000, builder="batchrule")],
        train=dict(name="chrT", n=2000, seed=12, depth=300.0, band=110, n_loops=500, loop_max=70),
        lower=6, upper=300, res=10000, w=5, weight="weight", min_prob=0.5,
        forest=dict(n_estimators=20, max_depth=8, seed=1), store_inputs=False, full_taps=False),
    # BASELINE.json configs[1] at full size: the benchmark's chromosome (bench.py workload c2, seed 1234) and
    # the benchmark's forest (bench_data/c2.pkl). Only checksums of the large taps are stored.
    "c2": dict(
        chroms=[dict(name="chr1", n=24900, seed=1234, depth=300.0, band=330)],
        lower=6, upper=300, res=10000, w=5, weight="weight", min_prob=0.5,
        forest=dict(pretrained="bench_data/c2"), store_inputs=False, full_taps=False, compact=True),
}


def build_chrom(spec):
    if spec.get("builder") == "batchrule":
        from tests.cases import make_batchrule_chromosome
        return make_batchrule_chromosome(spec["name"], spec["n"])
    kw = {k: v for k, v in spec.items() if k not in ("name", "n")}
    return synth.make_chromosome(spec["name"], spec["n"], **kw)


def sample_negatives(ch, n_neg, w, dmax, seed):
    """Random stored pixels with finite balanced value, w < d <= dmax, not planted."""
    rng = np.random.default_rng(seed)
    d = ch.bin2 - ch.bin1
    ok = (d > w) & (d <= dmax) & np.isfinite(ch.weights[ch.bin1] * ch.weights[ch.bin2])
    idx = np.nonzero(ok)[0]
    planted = set(map(tuple, ch.loops.tolist()))
    pick = rng.choice(idx, size=min(idx.size, 2 * n_neg), replace=False)
    out = []
    for i in pick:
        p = (int(ch.bin1[i]), int(ch.bin2[i]))
        if p not in planted:
            out.append(p)
        if len(out) == n_neg:
            break
    return out


def train_forest(case, buildmatrix=trainUtils.buildmatrix):
    """Forest fitted on the reference's own training features (trainUtils.py:12-44)."""
    tch = build_chrom(case["train"])
    path = os.path.join(tempfile.mkdtemp(), "train.pkcool")
    coolio.PKCool.write(path, [tch], case["res"])
    lib = coolio.Cooler(path)
    balance = False if case["weight"] == "raw" else case["weight"]
    M = utils.tocsr(lib.matrix(balance=balance, sparse=True).fetch(tch.name))
    w = case["w"]
    pos = [tuple(p) for p in tch.loops.tolist()]
    neg = sample_negatives(tch, len(pos), w, case["upper"], seed=case["train"]["seed"] + 7)
    fpos = buildmatrix(M, pos, w=w)
    fneg = buildmatrix(M, neg, w=w)
    fneg = fneg[:len(fpos)]
    X = np.r_[fpos + fneg]
    y = np.r_[[1] * len(fpos) + [0] * len(fneg)]
    f = case["forest"]
    model = RandomForestClassifier(n_estimators=f["n_estimators"], max_depth=f["max_depth"],
                                   max_features="sqrt", class_weight="balanced", criterion="gini",
                                   n_jobs=1, random_state=f["seed"])
    model.fit(X, y)
    return model, X.shape


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(name):
    case = CASES[name]
    print("== case", name)
    chroms = [build_chrom(s) for s in case["chroms"]]
    tmp = tempfile.mkdtemp()
    cool = os.path.join(tmp, name + ".pkcool")
    coolio.PKCool.write(cool, chroms, case["res"])

    if "pretrained" in case["forest"]:
        pkl = os.path.join(ROOT, case["forest"]["pretrained"] + ".pkl")
        model = joblib.load(pkl)
        ff = flatten_forest(model)
        print("   forest: %s, nodes %d" % (case["forest"]["pretrained"], ff.n_nodes))
    else:
        model, tshape = train_forest(case)
        pkl = os.path.join(HERE, name + ".pkl")
        joblib.dump(model, pkl, compress=("xz", 3))
        ff = flatten_forest(model)
        ff.save(os.path.join(HERE, name + "_forest.npz"))
        print("   forest: trained on", tshape, "nodes", ff.n_nodes)
    compact = bool(case.get("compact"))

    out = {}
    meta = dict(case=case, checksums={c.name: c.checksum() for c in chroms},
                forest_nodes=ff.n_nodes)
    w = case["w"]
    lib = coolio.Cooler(cool)
    for ch in chroms:
        if case["weight"] == "raw":
            M = utils.tocsr(lib.matrix(balance=False, sparse=True).fetch(ch.name))
            X = scoreUtils.Chromosome(M, model=model, raw_M=M, weights=None, cname="chr" + ch.name.lstrip("chr"),
                                      lower=case["lower"], upper=case["upper"], res=case["res"], width=w)
        else:
            M = utils.tocsr(lib.matrix(balance=case["weight"], sparse=True).fetch(ch.name))
            raw_M = utils.tocsr(lib.matrix(balance=False, sparse=True).fetch(ch.name))
            weights = lib.bins().fetch(ch.name)[case["weight"]].values
            X = scoreUtils.Chromosome(M, model=model, raw_M=raw_M, weights=weights,
                                      cname="chr" + ch.name.lstrip("chr"),
                                      lower=case["lower"], upper=case["upper"], res=case["res"], width=w)
        k = ch.name + "/"
        out[k + "exp_arr"] = np.asarray(X.exp_arr, dtype=np.float64)
        out[k + "background"] = np.asarray(X.background, dtype=np.float64)
        if not compact:
            out[k + "ridx"] = X.ridx.astype(np.int32)
            out[k + "cidx"] = X.cidx.astype(np.int32)
        coords = [(r, c) for r, c in zip(X.ridx, X.cidx)]
        # taps (iii)/(iv) over ALL candidates in one call (no batching quirk here)
        fea, clist = X.getwindow(coords) if coords else (np.zeros((0, (2 * w + 1) ** 2)), np.zeros((0, 2)))
        fea = np.asarray(fea, dtype=np.float64).reshape(-1, (2 * w + 1) ** 2)
        clist = np.asarray(clist).reshape(-1, 2)
        fea32 = fea.astype(np.float32)
        proba = model.predict_proba(fea)[:, 1] if fea.shape[0] else np.zeros(0)
        leaves = np.stack([e.apply(fea32) for e in model.estimators_], axis=1).astype(np.int32) \
            if fea.shape[0] else np.zeros((0, len(model.estimators_)), np.int32)
        if not compact:
            out[k + "clist"] = clist.astype(np.int32)
            out[k + "proba"] = proba
        meta.setdefault("sha", {})[ch.name] = dict(fea64=sha(fea), fea32=sha(fea32), leaves=sha(leaves),
                                                  n_windows=int(fea.shape[0]), n_candidates=int(X.ridx.size),
                                                  ridx=sha(X.ridx.astype(np.int32)), cidx=sha(X.cidx.astype(np.int32)),
                                                  clist=sha(clist.astype(np.int32)), proba=sha(proba))
        if case["full_taps"]:
            out[k + "fea64_head"] = fea[:128]      # float64 for the first 128 windows
            out[k + "fea32"] = fea32               # float32 (what the forest sees) for all
            out[k + "leaves"] = leaves
        if case["store_inputs"]:
            out[k + "bin1"], out[k + "bin2"], out[k + "count"] = ch.bin1, ch.bin2, ch.count
            out[k + "weights"] = ch.weights
        print("   %s: n=%d candidates=%d windows=%d" % (ch.name, ch.n, X.ridx.size, fea.shape[0]))

    # tap (v): the reference drivers end to end
    bed = os.path.join(tmp, name + ".bedpe")
    ns = argparse.Namespace(path=cool, model=pkl, output=bed, resolution=case["res"],
                            lower=case["lower"], upper=case["upper"], minimum_prob=case["min_prob"],
                            clr_weight_name=case["weight"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        if case.get("genome"):
            ns.chroms = case.get("chroms_arg", ["#", "X"])
            score_genome.main(ns)
        else:
            ns.chrom = chroms[0].name
            score_chromosome.main(ns)
    meta["stdout"] = buf.getvalue()
    with open(bed) as fh:
        bedtxt = fh.read()
    meta["bedpe_sha"] = hashlib.sha256(bedtxt.encode()).hexdigest()
    meta["bedpe_rows"] = bedtxt.count("\n")
    if compact:
        # the text is large: keep its checksum and the records (columns 2, 5 / res, 7, 8 of the text) instead
        rows = [ln.split("\t") for ln in bedtxt.splitlines()]
        out["records/x"] = np.array([int(r[1]) // case["res"] for r in rows], dtype=np.int32)
        out["records/y"] = np.array([int(r[4]) // case["res"] for r in rows], dtype=np.int32)
        out["records/prob"] = np.array([float(r[6]) for r in rows], dtype=np.float64)
        out["records/value"] = np.array([float(r[7]) for r in rows], dtype=np.float64)
    else:
        with open(os.path.join(HERE, name + ".bedpe"), "w") as fh:
            fh.write(bedtxt)
    print("   bedpe rows:", bedtxt.count("\n"))

    # tap (vi): pool
    for thr in (0.9, case["min_prob"]):
        pooled = os.path.join(tmp, "pool.bedpe")
        try:
            call_loops.main(argparse.Namespace(infile=bed, outfile=pooled, threshold=thr,
                                               resolution=case["res"]))
            with open(pooled) as fh:
                ptxt = fh.read()
        except Exception as e:  # the reference's clustering can fail on tiny inputs
            ptxt = "ERROR " + type(e).__name__
        if compact:
            meta.setdefault("pool_sha", {})[str(thr)] = hashlib.sha256(ptxt.encode()).hexdigest()
            meta.setdefault("pool_rows", {})[str(thr)] = ptxt.count("\n")
        else:
            with open(os.path.join(HERE, "%s.pool_t%s.bedpe" % (name, thr)), "w") as fh:
                fh.write(ptxt)
        print("   pool t=%s rows: %s" % (thr, ptxt.count("\n")))

    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n)
