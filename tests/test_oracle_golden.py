"""The oracle against outputs of the reference itself (tests/golden/, made by
make_golden.py from /root/reference): candidate set, expected curve, window
features, leaves, probabilities, bedpe text. CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import peakachu_oracle as po
from peakachu_b200 import coolio
from tests.cases import ALL_CASES, BIG_CASES, FULL_TAP_CASES, Case


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _chromosome(case, lib, ch, model):
    cfg = case.cfg
    cname = "chr" + ch.name.lstrip("chr")
    if cfg["weight"] == "raw":
        M = po.tocsr(lib.matrix(balance=False, sparse=True).fetch(ch.name))
        return po.Chromosome(M, model=model, raw_M=M, weights=None, cname=cname, lower=cfg["lower"],
                             upper=cfg["upper"], res=cfg["res"], width=cfg["w"])
    M = po.tocsr(lib.matrix(balance=cfg["weight"], sparse=True).fetch(ch.name))
    raw_M = po.tocsr(lib.matrix(balance=False, sparse=True).fetch(ch.name))
    weights = lib.bins().fetch(ch.name)[cfg["weight"]].values
    return po.Chromosome(M, model=model, raw_M=raw_M, weights=weights, cname=cname, lower=cfg["lower"],
                         upper=cfg["upper"], res=cfg["res"], width=cfg["w"])


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_taps_match_reference(name, tmp_path):
    case = Case(name)
    model = case.model()
    lib = coolio.Cooler(case.write_cool(tmp_path))
    for ch in case.chroms:
        X = _chromosome(case, lib, ch, model)
        k = ch.name + "/"
        assert np.array_equal(X.exp_arr, case.z[k + "exp_arr"])              # tap (ii), bit-exact
        assert np.array_equal(X.ridx, case.z[k + "ridx"])                    # tap (i)
        assert np.array_equal(X.cidx, case.z[k + "cidx"])
        fea, clist = X.getwindow(np.stack([X.ridx, X.cidx], axis=1))         # tap (iii)
        assert np.array_equal(clist, case.z[k + "clist"])
        sh = case.meta["sha"][ch.name]
        assert fea.shape[0] == sh["n_windows"]
        assert _sha(fea) == sh["fea64"]                                      # float64 bit-exact
        fea32 = fea.astype(np.float32)
        assert _sha(fea32) == sh["fea32"]
        leaves = po.forest_apply(case.forest, fea32)                         # tap (iv)
        assert _sha(leaves.astype(np.int32)) == sh["leaves"]
        proba = po.forest_proba(case.forest, fea32)
        assert np.array_equal(proba, case.z[k + "proba"])                    # bit-exact float64
        if name in FULL_TAP_CASES:
            assert np.array_equal(fea[:128], case.z[k + "fea64_head"])
            assert np.array_equal(fea32, case.z[k + "fea32"])
            assert np.array_equal(leaves, case.z[k + "leaves"])


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_bedpe_matches_reference(name, tmp_path):
    case = Case(name)
    cfg = case.cfg
    lib = coolio.Cooler(case.write_cool(tmp_path))
    out = os.path.join(str(tmp_path), "o.bedpe")
    names = [c.name for c in case.chroms]
    if cfg.get("genome"):
        from peakachu_b200.score_genome import select_chromosomes
        names = select_chromosomes(names, case.chroms_arg())                 # score_genome.py:39-44
    po.score_map(lib, case.model(), names, weight_name=cfg["weight"], lower=cfg["lower"],
                 upper=cfg["upper"], res=cfg["res"], min_prob=cfg["min_prob"], output=out,
                 genome=bool(cfg.get("genome")))
    assert open(out).read() == case.bedpe                                    # tap (v)


@pytest.mark.parametrize("name", BIG_CASES)
def test_oracle_full_size_matches_reference(name, tmp_path):
    """BASELINE configs[1] at full size (24,900 bins, the benchmark's map and forest): every tap of the
    oracle against checksums of the reference's own run (make_golden.py, compact case)."""
    case = Case(name)
    cfg, model = case.cfg, case.model()
    lib = coolio.Cooler(case.write_cool(tmp_path))
    ch = case.chroms[0]
    X = _chromosome(case, lib, ch, model)
    sh = case.meta["sha"][ch.name]
    assert np.array_equal(X.exp_arr, case.z[ch.name + "/exp_arr"])
    assert X.ridx.size == sh["n_candidates"]
    assert _sha(X.ridx.astype(np.int32)) == sh["ridx"] and _sha(X.cidx.astype(np.int32)) == sh["cidx"]
    fea, clist = X.getwindow(np.stack([X.ridx, X.cidx], axis=1))
    assert fea.shape[0] == sh["n_windows"] and _sha(clist.astype(np.int32)) == sh["clist"]
    assert _sha(fea) == sh["fea64"]
    fea32 = fea.astype(np.float32)
    assert _sha(fea32) == sh["fea32"]
    assert _sha(po.forest_proba(case.forest, fea32)) == sh["proba"]
    out = os.path.join(str(tmp_path), "o.bedpe")
    po.score_map(lib, model, [ch.name], weight_name=cfg["weight"], lower=cfg["lower"], upper=cfg["upper"],
                 res=cfg["res"], min_prob=cfg["min_prob"], output=out)
    txt = open(out).read()
    assert hashlib.sha256(txt.encode()).hexdigest() == case.meta["bedpe_sha"]
    assert txt.count("\n") == case.meta["bedpe_rows"] == case.z["records/x"].size


def test_batch_rule_fixture_drops_the_lone_window():
    """The fixture exists for scoreUtils.py:104-108: four batches keep 2 / 1 / 0 / 3 windows, the reference's
    bedpe holds the five records of the first and the last batch."""
    case = Case("batchrule")
    clist = case.z["chr7/clist"]
    assert clist.shape[0] == 6 and case.meta["sha"]["chr7"]["n_candidates"] > 3 * po.BATCH
    rank = {(int(x), int(y)): i for i, (x, y) in enumerate(zip(case.z["chr7/ridx"], case.z["chr7/cidx"]))}
    per_batch = np.bincount([rank[(int(x), int(y))] // po.BATCH for x, y in clist], minlength=4)
    assert per_batch.tolist() == [2, 1, 0, 3]
    assert case.bedpe.count("\n") == 5


@pytest.mark.parametrize("name", ["tiny", "tiny_raw", "w7"])
def test_oracle_buildmatrix_matches_reference(name, tmp_path):
    """trainUtils.buildmatrix (trainUtils.py:12-44): features of the reference itself
    (tests/golden/make_buildmatrix_golden.py), float64 bit-exact."""
    case = Case(name)
    cfg, ch = case.cfg, case.chroms[0]
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "buildmatrix.npz"))
    lib = coolio.Cooler(case.write_cool(tmp_path))
    balance = False if cfg["weight"] == "raw" else cfg["weight"]
    M = po.tocsr(lib.matrix(balance=balance, sparse=True).fetch(ch.name))
    fea = po.buildmatrix(M, [tuple(p) for p in g[name + "/coords"].tolist()], w=cfg["w"])
    assert np.array_equal(np.array(fea), g[name + "/fea"])
