"""Parity of the CUDA path (through the C ABI) against (a) the golden fixtures made
by the reference and (b) the numpy oracle on the same inputs. Bit-exact for the
expected curve, candidate set, float64/float32 features, leaf ids and
probabilities; byte-identical bedpe text."""
import argparse
import ctypes as C
import os

import numpy as np
import pytest

from tests.cases import ALL_CASES, BIG_CASES, FULL_TAP_CASES, Case

pytestmark = pytest.mark.gpu


def _gpu_chromosome(case, ch, forest=None):
    from peakachu_b200.scoreUtils import Chromosome
    cfg = case.cfg
    weights = None if cfg["weight"] == "raw" else ch.weights
    return Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, weights, ch.n, forest or case.forest,
                                  lower=cfg["lower"], upper=cfg["upper"], cname="chr" + ch.name.lstrip("chr"),
                                  res=cfg["res"], width=cfg["w"])


@pytest.mark.parametrize("name", ALL_CASES)
def test_taps_match_reference_golden(name):
    import torch
    from peakachu_b200 import _lib
    from peakachu_b200.scoreUtils import DeviceForest
    case = Case(name)
    df = DeviceForest.of(case.forest, 0)
    for ch in case.chroms:
        k = ch.name + "/"
        X = _gpu_chromosome(case, ch, df)
        assert np.array_equal(X.exp_arr, case.z[k + "exp_arr"])                  # (ii) bit-exact
        assert np.array_equal(X.ridx, case.z[k + "ridx"])                        # (i) set and order
        assert np.array_equal(X.cidx, case.z[k + "cidx"])
        keep, f32, f64 = X.window_features(want64=True)                          # (iii)
        clist = np.stack([X.ridx[keep], X.cidx[keep]], axis=1)
        assert np.array_equal(clist, case.z[k + "clist"])
        if name in FULL_TAP_CASES:
            assert np.array_equal(f64[keep][:128], case.z[k + "fea64_head"])     # float64 bit-exact
            assert np.array_equal(f32[keep], case.z[k + "fea32"])                # float32 identical
        import hashlib
        sh = case.meta["sha"][ch.name]
        assert hashlib.sha256(np.ascontiguousarray(f64[keep]).tobytes()).hexdigest() == sh["fea64"]
        assert hashlib.sha256(np.ascontiguousarray(f32[keep]).tobytes()).hexdigest() == sh["fea32"]
        # (iii) again, tapped INSIDE the product kernel: the float32 rows k_score_fused hands its forest walk
        keep_f, f32_f = X.fused_window_features()
        assert np.array_equal(keep_f, keep)
        assert hashlib.sha256(np.ascontiguousarray(f32_f[keep_f]).tobytes()).hexdigest() == sh["fea32"]
        assert not f32_f[~keep_f].any()
        if name in FULL_TAP_CASES:
            assert np.array_equal(f32_f[keep_f], case.z[k + "fea32"])
        # (iv) leaves + probabilities from the device forest on those rows
        xs = torch.from_numpy(np.ascontiguousarray(f32[keep])).cuda()
        leaves = torch.empty((xs.shape[0], case.forest.n_trees), dtype=torch.int32, device="cuda")
        proba = torch.empty(xs.shape[0], dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().pk_forest_apply(df.handle, C.c_void_p(xs.data_ptr()), xs.shape[0],
                                              C.c_void_p(leaves.data_ptr()), C.c_void_p(proba.data_ptr()), None))
        torch.cuda.synchronize()
        assert hashlib.sha256(leaves.cpu().numpy().tobytes()).hexdigest() == sh["leaves"]
        assert np.array_equal(proba.cpu().numpy(), case.z[k + "proba"])          # bit-exact float64
        if name in FULL_TAP_CASES:
            assert np.array_equal(leaves.cpu().numpy(), case.z[k + "leaves"])
        X.close()


def _windows_after_batch_rule(case, k):
    """(clist, proba) of the golden windows that survive scoreUtils.py:104-108: windows of a
    100,000-candidate batch that keeps at most one window are dropped by Chromosome.score."""
    clist, proba = case.z[k + "clist"], case.z[k + "proba"]
    rank = {(int(x), int(y)): i for i, (x, y) in enumerate(zip(case.z[k + "ridx"], case.z[k + "cidx"]))}
    batch = np.array([rank[(int(x), int(y))] // 100000 for x, y in clist], dtype=np.int64)
    ok = np.bincount(batch, minlength=1)[batch] > 1 if batch.size else np.zeros(0, bool)
    return clist[ok], proba[ok], int(clist.shape[0])


@pytest.fixture
def tuning():
    from peakachu_b200 import _lib

    def set_fused(v):
        _lib.check(_lib.lib().pk_set_tuning(b"fused", v))
    yield set_fused
    set_fused(-1)


@pytest.mark.parametrize("fused", [0, 1, 2])
@pytest.mark.parametrize("name", ALL_CASES)
def test_all_window_probabilities_bit_exact(name, fused, tuning):
    """Every kept window's probability (threshold below 0 keeps them all), through the
    separate kernels (fused=0) and both fused-kernel variants."""
    tuning(fused)
    case = Case(name)
    for ch in case.chroms:
        k = ch.name + "/"
        X = _gpu_chromosome(case, ch)
        x, y, p, v = X.score_records(-1.0)
        clist, proba, n_windows = _windows_after_batch_rule(case, k)
        order = np.lexsort((clist[:, 1], clist[:, 0]))
        assert np.array_equal(x, clist[order, 0]) and np.array_equal(y, clist[order, 1])
        assert np.array_equal(p, proba[order])
        assert X.n_windows == n_windows
        X.close()


@pytest.mark.parametrize("cf", [0, 1])
@pytest.mark.parametrize("name", ALL_CASES)
def test_both_forest_encodings_are_exact(name, cf):
    """The fused kernel walks the forest on either node encoding (own feature per node, or the
    children's features per node: pk_set_tuning("child_features")); both give the reference's
    probabilities bit for bit, with and without pruning."""
    from peakachu_b200 import _lib
    L = _lib.lib()
    case = Case(name)
    try:
        _lib.check(L.pk_set_tuning(b"child_features", cf))
        for thre in (-1.0, case.cfg["min_prob"]):
            for ch in case.chroms:
                k = ch.name + "/"
                X = _gpu_chromosome(case, ch)
                x, y, p, v = X.score_records(thre)
                clist, proba, _ = _windows_after_batch_rule(case, k)
                order = np.lexsort((clist[:, 1], clist[:, 0]))
                if thre < 0:
                    assert np.array_equal(x, clist[order, 0]) and np.array_equal(y, clist[order, 1])
                    assert np.array_equal(p, proba[order])
                else:               # a subset of the windows (the 100,000-candidate batch rule may drop more)
                    ref = {(int(a), int(b)): float(c) for a, b, c in zip(clist[:, 0], clist[:, 1], proba)}
                    assert all(ref[(int(a), int(b))] == float(c) and c > thre for a, b, c in zip(x, y, p))
                X.close()
    finally:
        _lib.check(L.pk_set_tuning(b"child_features", -1))


@pytest.mark.parametrize("fused", [0, 1, 2])
@pytest.mark.parametrize("name", ALL_CASES)
def test_bedpe_identical_to_reference(name, fused, tuning, tmp_path):
    from peakachu_b200 import score_chromosome, score_genome
    tuning(fused)
    case = Case(name)
    cfg = case.cfg
    cool = case.write_cool(tmp_path)
    out = os.path.join(str(tmp_path), "gpu.bedpe")
    ns = argparse.Namespace(path=cool, model=case.pkl, output=out, resolution=cfg["res"], lower=cfg["lower"],
                            upper=cfg["upper"], minimum_prob=cfg["min_prob"], clr_weight_name=cfg["weight"])
    if cfg.get("genome"):
        ns.chroms = case.chroms_arg()
        score_genome.main(ns)
    else:
        ns.chrom = case.chroms[0].name
        score_chromosome.main(ns)
    txt = open(out).read()
    assert txt == case.bedpe                                                     # (v)
    # (vi) `peakachu pool` stays on the host and reads this text; the reference's pooled loops of the
    # golden bedpe are rows of OUR output, field for field (peakacluster re-prints prob and value)
    rows = {tuple(ln.split("\t")[i] for i in (0, 1, 4)): ln.split("\t")[6:8] for ln in txt.splitlines()}
    for thr in (0.9, cfg["min_prob"]):
        for ln in case.pool(thr).splitlines():
            if ln.startswith("ERROR"):
                continue
            f = ln.split("\t")
            assert rows[(f[0], f[1], f[4])] == f[6:8]


@pytest.mark.parametrize("name", ["tiny", "tiny_raw", "genome", "gnames"])
def test_expected_fit_on_the_host(name, tmp_path, monkeypatch):
    """PEAKACHU_B200_EXPECTED=host: the curve comes from the installed scikit-learn (the reference's own
    calls, utils.py:173-176) on the device's per-distance means and goes back with pk_chrom_set_expected;
    score_genome then runs its units through scoreUtils.Chromosome instead of the engine. With the pinned
    versions installed both modes give the reference's bedpe."""
    from peakachu_b200 import _lib, score_chromosome, score_genome
    monkeypatch.setenv("PEAKACHU_B200_EXPECTED", "host")
    monkeypatch.setattr(_lib, "_expected_mode", None)
    assert _lib.expected_mode() == "host"
    case = Case(name)
    cfg = case.cfg
    cool = case.write_cool(tmp_path)
    out = os.path.join(str(tmp_path), "gpu.bedpe")
    ns = argparse.Namespace(path=cool, model=case.pkl, output=out, resolution=cfg["res"], lower=cfg["lower"],
                            upper=cfg["upper"], minimum_prob=cfg["min_prob"], clr_weight_name=cfg["weight"])
    if cfg.get("genome"):
        ns.chroms = case.chroms_arg()
        score_genome.main(ns)
    else:
        ns.chrom = case.chroms[0].name
        score_chromosome.main(ns)
    assert open(out).read() == case.bedpe
    monkeypatch.setattr(_lib, "_expected_mode", None)


@pytest.mark.parametrize("name", ["tiny", "tiny_raw", "lowdepth", "c1", "w7"])
def test_windows_as_tma_boxes(name, tmp_path):
    """pk_set_tuning("tma", 2): the fused kernel of both widths fetches every window as one TMA box from the
    row-major band copy (the default does so for w = 7 only). Float32 features inside the kernel (the tap) and
    the bedpe are the reference's."""
    from peakachu_b200 import _lib, score_chromosome
    L = _lib.lib()
    case = Case(name)
    cfg = case.cfg
    try:
        _lib.check(L.pk_set_tuning(b"tma", 2))
        import hashlib
        for ch in case.chroms:
            X = _gpu_chromosome(case, ch)
            keep, f32 = X.fused_window_features()
            assert hashlib.sha256(np.ascontiguousarray(f32[keep]).tobytes()).hexdigest() == case.meta["sha"][ch.name]["fea32"]
            if name in FULL_TAP_CASES:
                assert np.array_equal(f32[keep], case.z[ch.name + "/fea32"])
            X.close()
        cool = case.write_cool(tmp_path)
        out = os.path.join(str(tmp_path), "gpu.bedpe")
        score_chromosome.main(argparse.Namespace(path=cool, model=case.pkl, output=out, resolution=cfg["res"], lower=cfg["lower"],
                                                 upper=cfg["upper"], minimum_prob=cfg["min_prob"], clr_weight_name=cfg["weight"],
                                                 chrom=case.chroms[0].name))
        assert open(out).read() == case.bedpe
    finally:
        _lib.check(L.pk_set_tuning(b"tma", 1))


@pytest.mark.parametrize("name,group", [("tiny", ""), ("tiny_raw", ""), ("genome", "resolutions/10000")])
def test_bedpe_from_a_real_cool_file(name, group, tmp_path):
    """the same CLI on an HDF5 cooler file (.cool / .mcool::group, read by
    h5mini) writes the reference's bedpe; inter-chromosomal pixels in the file are ignored."""
    from peakachu_b200 import score_chromosome, score_genome
    from tests import h5write
    case = Case(name)
    cfg = case.cfg
    nb = np.array([c.n for c in case.chroms])
    off = np.concatenate([[0], np.cumsum(nb)])
    trans = [(off[0] + 3, off[1] + 5, 7), (off[0] + nb[0] - 1, off[-1] - 1, 2)] if len(nb) > 1 else None
    cool = os.path.join(str(tmp_path), name + (".mcool" if group else ".cool"))
    h5write.write_cool(cool, case.chroms, cfg["res"], trans=trans, group=group, chunk=5000)
    out = os.path.join(str(tmp_path), "gpu.bedpe")
    ns = argparse.Namespace(path=cool + ("::/" + group if group else ""), model=case.pkl, output=out,
                            resolution=cfg["res"], lower=cfg["lower"], upper=cfg["upper"],
                            minimum_prob=cfg["min_prob"], clr_weight_name=cfg["weight"])
    if cfg.get("genome"):
        ns.chroms = ["#", "X"]
        score_genome.main(ns)
    else:
        ns.chrom = case.chroms[0].name
        score_chromosome.main(ns)
    assert open(out).read() == case.bedpe


@pytest.mark.parametrize("container", ["pkcool", "cool"])
def test_divisive_weight_column(container, tmp_path):
    """A weight column flagged `divisive_weights` (hic2cool's KR / VC): cooler's matrix(balance=name) divides by
    the weights, while the reference hands the raw column to Chromosome, which uses it in the Poisson filter
    (score_chromosome.py:42-44, scoreUtils.py:55-57). The CLI reproduces the oracle run on such a map -- and
    differs from a run that treats the column as multiplicative."""
    import copy
    from oracle import peakachu_oracle as po
    from peakachu_b200 import coolio, score_chromosome
    from tests import h5write
    case = Case("tiny")
    cfg = case.cfg
    ch = copy.copy(case.chroms[0])
    with np.errstate(divide="ignore", invalid="ignore"):
        kr = 1.0 / ch.weights                       # the column as hic2cool stores it: balanced = count / (kr_i kr_j)
    ref_path = os.path.join(str(tmp_path), "ref.pkcool")
    ch_kr = copy.copy(ch)
    ch_kr.weights = kr
    coolio.PKCool.write(ref_path, [ch_kr], cfg["res"], weight_name="KR", divisive=True)
    lib = coolio.Cooler(ref_path)                  # the stand-in inverts a divisive column like cooler does
    out_o = os.path.join(str(tmp_path), "oracle.bedpe")
    po.score_map(lib, case.model(), [ch.name], weight_name="KR", lower=cfg["lower"], upper=cfg["upper"],
                 res=cfg["res"], min_prob=cfg["min_prob"], output=out_o)
    if container == "pkcool":
        path = ref_path
    else:
        path = os.path.join(str(tmp_path), "kr.cool")
        h5write.write_cool(path, [ch], cfg["res"], extra_bins={"KR": kr})
    out = os.path.join(str(tmp_path), "gpu.bedpe")
    ns = argparse.Namespace(path=path, model=case.pkl, output=out, resolution=cfg["res"], lower=cfg["lower"],
                            upper=cfg["upper"], minimum_prob=cfg["min_prob"], clr_weight_name="KR", chrom=ch.name)
    score_chromosome.main(ns)
    assert open(out).read() == open(out_o).read()
    assert open(out).read() != case.bedpe           # the Poisson filter saw the raw column, not its reciprocal


def test_forest_npz_and_pkl_give_same_tables():
    from peakachu_b200.forest import load_model
    case = Case("tiny")
    a, _ = load_model(case.pkl)
    b = case.forest
    for f in ("node_offset", "feature", "threshold", "left", "right", "missing_left", "leaf_p1"):
        assert np.array_equal(getattr(a, f), getattr(b, f))


def test_dropin_constructor_with_scipy_matrices(tmp_path):
    """The reference-shaped constructor Chromosome(M, model, raw_M, weights, ...)."""
    from oracle import peakachu_oracle as po
    from peakachu_b200 import coolio
    from peakachu_b200.scoreUtils import Chromosome
    case = Case("tiny")
    cfg = case.cfg
    lib = coolio.Cooler(case.write_cool(tmp_path))
    ch = case.chroms[0]
    M = po.tocsr(lib.matrix(balance="weight", sparse=True).fetch(ch.name))
    raw_M = po.tocsr(lib.matrix(balance=False, sparse=True).fetch(ch.name))
    weights = lib.bins().fetch(ch.name)["weight"].values
    X = Chromosome(M, model=case.model(), raw_M=raw_M, weights=weights, cname="chr1", lower=cfg["lower"],
                   upper=cfg["upper"], res=cfg["res"], width=cfg["w"])
    prob, val = X.score(thre=cfg["min_prob"])
    O = po.Chromosome(M, model=case.model(), raw_M=raw_M, weights=weights, cname="chr1", lower=cfg["lower"],
                      upper=cfg["upper"], res=cfg["res"], width=cfg["w"])
    oprob, oval = O.score(thre=cfg["min_prob"])
    assert (prob != oprob).nnz == 0 and (val != oval).nnz == 0
    a, b = os.path.join(str(tmp_path), "a"), os.path.join(str(tmp_path), "b")
    X.writeBed(a, prob, val)
    O.writeBed(b, oprob, oval)
    assert open(a).read() == open(b).read() == case.bedpe


def test_row_tiles_equal_whole_chromosome():
    """Band row tiles (the multi-GPU seam) reproduce the whole-chromosome records."""
    from peakachu_b200 import _lib, shard
    from peakachu_b200.scoreUtils import DeviceForest

    class OneChrom:
        def __init__(self, ch): self.ch = ch
        def upper_pixels(self, k): return self.ch.bin1, self.ch.bin2, self.ch.count
        def weights(self, k, name): return self.ch.weights
        def nbins(self, k): return self.ch.n

    case = Case("c1")
    cfg = case.cfg
    ch = case.chroms[0]
    lib = OneChrom(ch)
    kw = dict(correct="weight", lower=cfg["lower"], upper=cfg["upper"], res=cfg["res"], device=0,
              min_prob=cfg["min_prob"])
    whole = shard.score_units(lib, [("chr1", 0, ch.n)], case.forest, **kw)
    tiles = shard.score_units(lib, [("chr1", 0, 700), ("chr1", 700, 1301), ("chr1", 1301, ch.n)], case.forest, **kw)
    tw = shard.assemble_text(["chr1"], [whole], cfg["res"])["chr1"]
    tt = shard.assemble_text(["chr1"], [tiles], cfg["res"])["chr1"]
    assert tw == tt == case.bedpe


def test_errors_are_loud():
    from peakachu_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.pk_chrom_create(0, 0, 5, 6, 300, 1, None, C.byref(h)) != 0
    assert b"n_bins" in L.pk_last_error()
    _lib.check(L.pk_chrom_create(0, 500, 5, 6, 300, 1, None, C.byref(h)))
    n = C.c_int64()
    assert L.pk_chrom_find_candidates(h, 0, 500, C.byref(n)) == -4         # PK_ESTATE: nothing uploaded
    L.pk_chrom_destroy(h)


def test_upload_paths_agree():
    """Order-free scatter path, sorted (tiled) path and direct CSR upload build the same
    band: identical expected curve, candidates and records; a false PK_PIXELS_SORTED
    promise is reported."""
    from peakachu_b200 import _lib
    from peakachu_b200.scoreUtils import Chromosome
    case = Case("tiny")
    cfg = case.cfg
    ch = case.chroms[0]
    kw = dict(lower=cfg["lower"], upper=cfg["upper"], cname="chr1", res=cfg["res"], width=cfg["w"])
    A = Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, ch.weights, ch.n, case.forest, sorted_pixels=True, **kw)
    rng = np.random.default_rng(0)
    perm = rng.permutation(ch.bin1.size)
    B = Chromosome.from_pixels(ch.bin1[perm], ch.bin2[perm], ch.count[perm], ch.weights, ch.n, case.forest,
                               sorted_pixels=False, **kw)
    assert np.array_equal(A.exp_arr, B.exp_arr) and np.array_equal(A.exp_arr, case.z["chr1/exp_arr"])
    assert np.array_equal(A.ridx, B.ridx) and np.array_equal(A.cidx, B.cidx)
    ra, rb = A.score_records(cfg["min_prob"]), B.score_records(cfg["min_prob"])
    assert all(np.array_equal(u, v) for u, v in zip(ra, rb))
    # CSR upload through the C ABI
    L = _lib.lib()
    h = C.c_void_p()
    _lib.check(L.pk_chrom_create(0, ch.n, cfg["w"], cfg["lower"], cfg["upper"], 1, None, C.byref(h)))
    rowptr = np.searchsorted(ch.bin1, np.arange(ch.n + 1)).astype(np.int64)
    b2, cnt, w = (np.ascontiguousarray(a) for a in (ch.bin2, ch.count, ch.weights))
    _lib.check(L.pk_chrom_upload_csr(h, _lib.ptr(rowptr), _lib.ptr(b2), _lib.ptr(cnt), b2.size, _lib.ptr(w), _lib.PK_MEM_HOST))
    _lib.check(L.pk_chrom_fit_expected(h))
    e = np.zeros(A.exp_arr.size)
    _lib.check(L.pk_chrom_get_expected(h, _lib.ptr(e, _lib.c_f64p)))
    assert np.array_equal(e, A.exp_arr)
    L.pk_chrom_destroy(h)
    # narrow columns (uint16 bin2 - bin1, uint16 count): same band, same records
    for name in ("tiny", "tiny_raw", "w7", "lowdepth"):
        cs = Case(name)
        c2, chx = cs.cfg, cs.chroms[0]
        kw2 = dict(lower=c2["lower"], upper=c2["upper"], cname=chx.name, res=c2["res"], width=c2["w"])
        rp = np.searchsorted(chx.bin1, np.arange(chx.n + 1)).astype(np.int64)
        assert int((chx.bin2 - chx.bin1).max()) <= 65535 and int(chx.count.max()) <= 65535
        wts = None if c2["weight"] == "raw" else chx.weights
        N = Chromosome.from_csr16(rp, (chx.bin2 - chx.bin1).astype(np.uint16), chx.count.astype(np.uint16), wts,
                                  chx.n, cs.forest, **kw2)
        W = Chromosome.from_csr(rp, chx.bin2, chx.count, wts, chx.n, cs.forest, **kw2)
        assert np.array_equal(N.exp_arr, W.exp_arr) and np.array_equal(N.exp_arr, cs.z[chx.name + "/exp_arr"])
        assert np.array_equal(N.ridx, W.ridx) and np.array_equal(N.cidx, W.cidx)
        rn, rw = N.score_records(c2["min_prob"]), W.score_records(c2["min_prob"])
        assert all(np.array_equal(u, v) for u, v in zip(rn, rw))
        # packed pixel rows (pk_chrom_upload_rows): exactly the band's distances, and far more than needed
        from peakachu_b200 import rowpack
        for nd_enc in (N._exp_len, N._exp_len + 70):
            P = Chromosome.from_rows(rowpack.pack_rows(rp, chx.bin2, chx.count, chx.n, nd_enc), wts, chx.n,
                                     cs.forest, **kw2)
            assert np.array_equal(P.exp_arr, W.exp_arr)
            assert np.array_equal(P.ridx, W.ridx) and np.array_equal(P.cidx, W.cidx)
            assert all(np.array_equal(u, v) for u, v in zip(P.score_records(c2["min_prob"]), rw))
            P.close()
        with pytest.raises(_lib.PKError, match="distances"):
            Chromosome.from_rows(rowpack.pack_rows(rp, chx.bin2, chx.count, chx.n, N._exp_len - 1), wts, chx.n,
                                 cs.forest, **kw2)
        N.close(); W.close()
    with pytest.raises(TypeError):
        Chromosome.from_csr16(rowptr, ch.bin2, ch.count, ch.weights, ch.n, case.forest, **kw)
    # a broken promise
    with pytest.raises(_lib.PKError, match="not sorted"):
        X = Chromosome.from_pixels(ch.bin1[perm], ch.bin2[perm], ch.count[perm], ch.weights, ch.n, case.forest,
                                   sorted_pixels=True, **kw)
        X.exp_arr
    A.close(); B.close()


def test_candidate_buffer_overflow_is_replayed():
    """min_prob below every probability and a map where far more than 1/8 of the band
    passes the Poisson filter: the device-side capacity flag triggers a transparent redo."""
    from peakachu_b200 import synth
    from peakachu_b200.scoreUtils import Chromosome
    from oracle import peakachu_oracle as po
    case = Case("tiny")
    rng = np.random.default_rng(3)
    n = 300
    # expected ~1 everywhere, but a third of the band carries a large count
    xs, ds = np.meshgrid(np.arange(n), np.arange(0, 80), indexing="ij")
    ok = xs + ds < n
    b1, b2 = xs[ok].astype(np.int32), (xs + ds)[ok].astype(np.int32)
    cnt = np.where(rng.random(b1.size) < 0.35, 40, 1).astype(np.int32)
    order = np.lexsort((b2, b1))
    b1, b2, cnt = b1[order], b2[order], cnt[order]
    w = np.full(n, 1.0)
    X = Chromosome.from_pixels(b1, b2, cnt, w, n, case.forest, lower=6, upper=60, cname="chr1", res=10000, width=5,
                               sorted_pixels=True)
    x, y, p, v = X.score_records(-1.0)
    ch = synth.SynthChrom("chr1", n, b1, b2, cnt, w, np.zeros((0, 2), np.int64))
    import tempfile, os
    from peakachu_b200 import coolio
    path = os.path.join(tempfile.mkdtemp(), "o.pkcool")
    coolio.PKCool.write(path, [ch], 10000)
    lib = coolio.Cooler(path)
    M = po.tocsr(lib.matrix(balance="weight", sparse=True).fetch("chr1"))
    raw = po.tocsr(lib.matrix(balance=False, sparse=True).fetch("chr1"))
    O = po.Chromosome(M, model=case.model(), raw_M=raw, weights=w, lower=6, upper=60, cname="chr1", res=10000, width=5)
    assert X.n_candidates == O.ridx.size and X.n_candidates > po.band_pixels(n, 6, 60, 5) // 8 + 4096 \
        if hasattr(po, "band_pixels") else X.n_candidates == O.ridx.size
    assert np.array_equal(X.ridx, O.ridx) and np.array_equal(X.cidx, O.cidx)
    fea, clist = O.getwindow(np.stack([O.ridx, O.cidx], axis=1))
    proba = po.forest_proba(case.forest, fea.astype(np.float32))
    order = np.lexsort((clist[:, 1], clist[:, 0]))
    assert np.array_equal(x, clist[order, 0]) and np.array_equal(y, clist[order, 1]) and np.array_equal(p, proba[order])
    X.close()


def test_reciprocal_division_is_ieee_exact():
    """pk_div_r (two FMA corrections of a*RN(1/b)) against __ddiv_rn on 2e9 operand pairs,
    including quotient 1, near-exact quotients, all-ones significands and a in (b/2, b)."""
    from peakachu_b200 import _lib
    bad = C.c_int64(-1)
    _lib.check(_lib.lib().pk_selftest_divide(0, 2_000_000_000, 12345, C.byref(bad)))
    assert bad.value == 0


def test_c4_shape_matches_oracle(tmp_path):
    """BASELINE config 4 shape at reduced length: 5 kb, upper=600, w=7 (15x15 windows),
    the 200-tree bench forest; records bit-exact against the oracle."""
    import io
    from contextlib import redirect_stdout
    import joblib
    from oracle import peakachu_oracle as po
    from peakachu_b200 import coolio, synth
    from peakachu_b200.forest import FlatForest
    from peakachu_b200.scoreUtils import Chromosome
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    flat = FlatForest.load(os.path.join(root, "bench_data", "c4_forest.npz"))
    model = joblib.load(os.path.join(root, "bench_data", "c4.pkl"))
    ch = synth.make_chromosome("chr1", 2200, seed=77, depth=300.0, band=640, n_loops=120, loop_max=550)
    X = Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, ch.weights, ch.n, flat, lower=6, upper=600, cname="chr1",
                               res=5000, width=7, sorted_pixels=True)
    x, y, p, v = X.score_records(0.5)
    path = os.path.join(str(tmp_path), "c4.pkcool")
    coolio.PKCool.write(path, [ch], 5000)
    lib = coolio.Cooler(path)
    M = po.tocsr(lib.matrix(balance="weight", sparse=True).fetch("chr1"))
    raw = po.tocsr(lib.matrix(balance=False, sparse=True).fetch("chr1"))
    O = po.Chromosome(M, model=model, raw_M=raw, weights=ch.weights, lower=6, upper=600, cname="chr1", res=5000, width=7)
    assert np.array_equal(X.exp_arr, O.exp_arr)
    assert np.array_equal(X.ridx, O.ridx) and np.array_equal(X.cidx, O.cidx)
    with redirect_stdout(io.StringIO()):
        prob, val = O.score(0.5)
    r, c = prob.nonzero()
    assert np.array_equal(x, r) and np.array_equal(y, c)
    assert np.array_equal(p, np.asarray(prob[r, c]).ravel()) and np.array_equal(v, np.asarray(val[r, c]).ravel())
    X.close()


def test_wide_band_takes_the_two_kernel_path(tmp_path):
    """upper = 1500 needs more shared memory for the expected curve than the fused kernel has
    left: the library must fall back to its separate feature and forest kernels, with the same
    records as the oracle."""
    import io
    from contextlib import redirect_stdout
    from oracle import peakachu_oracle as po
    from peakachu_b200 import coolio, synth
    from peakachu_b200.scoreUtils import Chromosome
    case = Case("tiny")
    ch = synth.make_chromosome("chr1", 2000, seed=5, depth=120.0, band=1530, n_loops=60, loop_max=1400)
    X = Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, ch.weights, ch.n, case.forest, lower=6, upper=1500,
                               cname="chr1", res=10000, width=5, sorted_pixels=True)
    x, y, p, v = X.score_records(0.5)
    assert X.stage_ms()["forest"] > 0.005        # the separate forest kernel ran
    path = os.path.join(str(tmp_path), "wide.pkcool")
    coolio.PKCool.write(path, [ch], 10000)
    lib = coolio.Cooler(path)
    M = po.tocsr(lib.matrix(balance="weight", sparse=True).fetch("chr1"))
    raw = po.tocsr(lib.matrix(balance=False, sparse=True).fetch("chr1"))
    O = po.Chromosome(M, model=case.model(), raw_M=raw, weights=ch.weights, lower=6, upper=1500, cname="chr1",
                      res=10000, width=5)
    assert np.array_equal(X.exp_arr, O.exp_arr)
    with redirect_stdout(io.StringIO()):
        prob, val = O.score(0.5)
    r, c = prob.nonzero()
    assert np.array_equal(x, r) and np.array_equal(y, c)
    assert np.array_equal(p, np.asarray(prob[r, c]).ravel()) and np.array_equal(v, np.asarray(val[r, c]).ravel())
    X.close()


def _random_forest_tables(sizes, n_features, seed):
    """FlatForest of randomly grown trees with the given node counts (odd)."""
    from peakachu_b200.forest import FlatForest
    rng = np.random.default_rng(seed)
    feats, thrs, lefts, rights, p1s, offs = [], [], [], [], [], [0]
    for n_nodes in sizes:
        left = np.full(n_nodes, -1, np.int32)
        right = np.full(n_nodes, -1, np.int32)
        leaves, used = [0], 1
        while used + 2 <= n_nodes:
            v = leaves.pop(int(rng.integers(0, len(leaves))))
            left[v], right[v] = used, used + 1
            leaves += [used, used + 1]
            used += 2
        internal = left >= 0
        feats.append(np.where(internal, rng.integers(0, n_features, n_nodes), -2).astype(np.int32))
        thrs.append(np.where(internal, rng.random(n_nodes), -2.0))
        lefts.append(left); rights.append(right)
        p1s.append(rng.integers(0, 8, n_nodes) / 7.0)
        offs.append(offs[-1] + n_nodes)
    n = offs[-1]
    return FlatForest(n_trees=len(sizes), n_features=n_features, node_offset=np.asarray(offs, np.int64),
                      feature=np.concatenate(feats), threshold=np.concatenate(thrs), left=np.concatenate(lefts),
                      right=np.concatenate(rights), missing_left=np.zeros(n, np.uint8), leaf_p1=np.concatenate(p1s))


@pytest.mark.parametrize("sizes", [[301, 4401, 5, 1, 4301, 77, 1201, 3, 2001, 601, 15, 4223],     # trees larger than a staging buffer
                                   [101, 60001, 33, 501]])                                         # right offsets beyond the fused encodings: separate kernels
def test_fused_walk_on_odd_forests(sizes, tuning):
    """Trees that do not fit the fused kernel's staging buffer (their tail is read from L2), single-leaf
    trees, groups of uneven size: the fused kernel on both node encodings equals the separate kernels."""
    from peakachu_b200 import _lib
    L = _lib.lib()
    case = Case("tiny")
    forest = _random_forest_tables(sizes, case.forest.n_features, seed=len(sizes))
    got = {}
    try:
        for key, fused, cf in (("separate", 0, 1), ("own", -1, 0), ("child", -1, 1)):
            tuning(fused)
            _lib.check(L.pk_set_tuning(b"child_features", cf))
            X = _gpu_chromosome(case, case.chroms[0], forest=forest)
            got[key] = X.score_records(-1.0)
            X.close()
    finally:
        _lib.check(L.pk_set_tuning(b"child_features", -1))
    assert got["separate"][0].size > 500
    for key in ("own", "child"):
        for a, b in zip(got["separate"], got[key]):
            assert np.array_equal(a, b), key


def test_fused_walk_with_nan_windows():
    """Constant windows scale to 0/0 = NaN features (utils.py:204-209) and sklearn routes those by
    missing_go_to_left. Half of this raw-mode chromosome is a constant plateau (with the expected curve
    set to 1 every window there is constant), the other half is noise, so warps mix NaN and ordinary
    pixels; the fused kernel on both encodings must equal the separate kernels, whose forest walk is
    pinned against sklearn on NaN rows (test_forest_nan_features_follow_missing_go_to_left)."""
    from peakachu_b200 import _lib
    from peakachu_b200.scoreUtils import Chromosome
    L = _lib.lib()
    case = Case("lowdepth")          # its forest was fitted with missing_go_to_left set on some nodes
    n, w, lower, upper = 400, 5, 6, 60
    rng = np.random.default_rng(4)
    b1, b2 = np.nonzero(np.triu(np.ones((n, n), bool)) & ~np.triu(np.ones((n, n), bool), upper + 2 * w + 1))
    cnt = np.where(b2 < n // 2, 3, rng.integers(1, 9, b1.size)).astype(np.int32)
    got = {}
    try:
        for key, fused, cf in (("separate", 0, -1), ("own", -1, 0), ("child", -1, 1)):
            _lib.check(L.pk_set_tuning(b"fused", fused))
            _lib.check(L.pk_set_tuning(b"child_features", cf))
            X = Chromosome.from_pixels(b1.astype(np.int32), b2.astype(np.int32), cnt, None, n, case.forest, lower=lower,
                                       upper=upper, cname="chr1", res=10000, width=w)
            el = X._exp_len
            exp, bg = np.ones(el), np.full(el, 0.01)
            _lib.check(L.pk_chrom_set_expected(X._h, exp.ctypes.data_as(_lib.c_f64p), bg.ctypes.data_as(_lib.c_f64p)))
            _lib.check(L.pk_chrom_find_candidates(X._h, 0, n, None))
            X._ncand = X._cand = None
            got[key] = X.score_records(-1.0)
            X.close()
    finally:
        _lib.check(L.pk_set_tuning(b"fused", -1))
        _lib.check(L.pk_set_tuning(b"child_features", -1))
    x, y, p, v = got["separate"]
    # (at distance `upper` the window's far corner lies on the first diagonal the band trim drops, scoreUtils.py:31)
    plateau = (y < n // 2 - w) & (y - x < upper)
    assert plateau.sum() > 1000 and (~plateau).sum() > 1000
    assert np.unique(p[plateau]).size == 1           # every constant window takes the all-NaN route
    for key in ("own", "child"):
        for a, b in zip(got["separate"], got[key]):
            assert np.array_equal(a, b), key


@pytest.mark.parametrize("w,case_name,combos", [
    (5, "lowdepth", [(0, -1), (2, -1), (2, 9), (2, 8)]),          # gather | TMA boxes | boxes a take ahead | a warp per group
    (7, "w7", [(0, -1), (1, -1), (1, 18), (1, 9)])])              # gather | default (boxes a take ahead, warp groups) | one group | 4 trees per round
def test_fused_kernels_on_hostile_weights(w, case_name, combos):
    """Every fused-kernel arrangement against the separate feature / forest kernels (pinned to the reference by the
    golden tests) on a balanced map whose weight column holds NaN, zero, negative, infinite, huge and tiny entries:
    the TMA kernels decide per window column whether (w_r w_c) count needs a finiteness test at all, from the
    largest high word of the weight products. Float32 features inside the kernel (bit patterns, NaN included) and
    the probabilities of all windows must be identical."""
    from peakachu_b200 import _lib
    from peakachu_b200.scoreUtils import Chromosome
    L = _lib.lib()
    case = Case(case_name)
    n, lower, upper = 520, 6, 70
    rng = np.random.default_rng(100 + w)
    b1, b2 = np.nonzero(np.triu(np.ones((n, n), bool)) & ~np.triu(np.ones((n, n), bool), upper + 2 * w + 1))
    keep = rng.random(b1.size) < 0.8
    b1, b2 = b1[keep].astype(np.int32), b2[keep].astype(np.int32)
    cnt = rng.integers(1, 30, b1.size).astype(np.int32)
    weights = np.exp(rng.normal(0.0, 0.3, n))
    special = rng.random(n)
    weights[special < 0.05] = np.nan
    weights[(special >= 0.05) & (special < 0.07)] = 0.0
    weights[(special >= 0.07) & (special < 0.08)] = -1.3
    weights[(special >= 0.08) & (special < 0.09)] = 1e160
    weights[(special >= 0.09) & (special < 0.10)] = 1e-170
    weights[(special >= 0.10) & (special < 0.105)] = np.inf
    weights[(special >= 0.105) & (special < 0.11)] = 1e-320           # denormal
    got = {}
    try:
        for tma, fused in [(0, 0)] + combos:
            _lib.check(L.pk_set_tuning(b"tma", tma))
            _lib.check(L.pk_set_tuning(b"fused", fused))
            X = Chromosome.from_pixels(b1, b2, cnt, weights, n, case.forest, lower=lower, upper=upper, cname="chr1",
                                       res=10000, width=w, sorted_pixels=True)
            el = X._exp_len
            exp = np.ascontiguousarray(5.0 / (1.0 + np.arange(el)) + 0.05)
            _lib.check(L.pk_chrom_set_expected(X._h, exp.ctypes.data_as(_lib.c_f64p), exp.ctypes.data_as(_lib.c_f64p)))
            _lib.check(L.pk_chrom_find_candidates(X._h, 0, n, None))
            X._ncand = X._cand = None
            if fused == 0:
                k0, f0 = X.window_features()[:2]
                got["tap"] = (k0, f0[k0].view(np.uint32).copy())
            else:
                kf, ff = X.fused_window_features()
                assert np.array_equal(kf, got["tap"][0]), (tma, fused)
                assert np.array_equal(ff[kf].view(np.uint32), got["tap"][1]), (tma, fused)
            got[(tma, fused)] = X.score_records(-1.0)
            X.close()
    finally:
        _lib.check(L.pk_set_tuning(b"fused", -1))
        _lib.check(L.pk_set_tuning(b"tma", 1))
    base = got[(0, 0)]
    assert base[0].size > 2000
    odd = ~np.isfinite(weights) | (weights <= 0) | (weights > 1e100) | (weights < 1e-100)
    near = np.convolve(odd.astype(int), np.ones(2 * w + 1, int), mode="same") > 0
    assert (near[base[0]] | near[base[1]]).sum() > 500            # windows that hold pixels of such bins: the slow path ran
    for key in combos:
        for a, b in zip(base, got[key]):
            assert np.array_equal(a, b, equal_nan=True), key


def test_forest_nan_features_follow_missing_go_to_left():
    """sklearn routes NaN features by missing_go_to_left; the forest tap must
    give the same leaves and probabilities on rows with NaNs."""
    import torch
    from peakachu_b200 import _lib
    from peakachu_b200.scoreUtils import DeviceForest
    case = Case("lowdepth")
    model = case.model()
    rng = np.random.default_rng(9)
    X = rng.random((500, case.forest.n_features)).astype(np.float32)
    X[rng.random(X.shape) < 0.15] = np.nan
    X[:5] = np.nan
    df = DeviceForest.of(case.forest, 0)
    xs = torch.from_numpy(X).cuda()
    leaves = torch.empty((X.shape[0], case.forest.n_trees), dtype=torch.int32, device="cuda")
    proba = torch.empty(X.shape[0], dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().pk_forest_apply(df.handle, C.c_void_p(xs.data_ptr()), X.shape[0],
                                          C.c_void_p(leaves.data_ptr()), C.c_void_p(proba.data_ptr()), None))
    torch.cuda.synchronize()
    want_leaves = np.stack([e.apply(X) for e in model.estimators_], axis=1)
    assert np.array_equal(leaves.cpu().numpy(), want_leaves)
    assert np.array_equal(proba.cpu().numpy(), model.predict_proba(X)[:, 1])


def test_edge_cases_empty_and_tiny():
    """No pixels -> the reference raises in IsotonicRegression.fit, we raise PKError; an upper
    bound below the lower bound -> no candidates, empty output; min_prob 1.0 -> no records."""
    from peakachu_b200 import _lib
    from peakachu_b200.scoreUtils import Chromosome
    case = Case("tiny")
    ch = case.chroms[0]
    e = np.zeros(0, np.int32)
    with pytest.raises(_lib.PKError, match="positive mean"):
        X = Chromosome.from_pixels(e, e, e, np.ones(100), 100, case.forest, lower=6, upper=60, width=5,
                                   sorted_pixels=True)
        X.exp_arr
    with pytest.raises(_lib.PKError, match="positive mean"):      # 4 bins: every diagonal has <= 10 entries
        one = np.array([1], np.int32)
        X = Chromosome.from_pixels(one * 0, one, one, np.ones(4), 4, case.forest, lower=6, upper=2, width=5,
                                   sorted_pixels=True)
        X.score_records(0.5)
    # upper (clamped to n - 2w) below lower: nothing to scan
    X = Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, ch.weights, ch.n, case.forest, lower=50, upper=20,
                               width=5, sorted_pixels=True)
    assert X.n_candidates == 0
    x, y, p, v = X.score_records(0.5)
    assert x.size == 0
    X.close()
    X = Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, ch.weights, ch.n, case.forest, lower=6, upper=60,
                               width=5, sorted_pixels=True)
    assert X.score_records(1.0)[0].size == 0          # probabilities never exceed 1
    prob, val = X.score(thre=1.0)
    assert prob.nnz == 0
    X.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny", "tiny_raw", "w7"])
def test_buildmatrix_matches_reference(name, tmp_path):
    """Training-set features (trainUtils.buildmatrix) from the CUDA feature
    kernel: float64 bit-exact against the reference's own output, through both the pixel-column
    entry point and the drop-in signature with scipy matrices."""
    from peakachu_b200 import coolio, trainUtils
    case = Case(name)
    cfg, ch = case.cfg, case.chroms[0]
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "buildmatrix.npz"))
    coords = [tuple(p) for p in g[name + "/coords"].tolist()]
    raw = cfg["weight"] == "raw"
    fea = trainUtils.buildmatrix_from_pixels(ch.bin1, ch.bin2, ch.count, None if raw else ch.weights, ch.n, coords,
                                             w=cfg["w"])
    assert np.array_equal(np.array(fea), g[name + "/fea"])
    lib = coolio.Cooler(case.write_cool(tmp_path))
    raw_M = lib.matrix(balance=False, sparse=True).fetch(ch.name).tocsr()
    if raw:
        fea2 = trainUtils.buildmatrix(raw_M, coords, w=cfg["w"])
    else:
        M = lib.matrix(balance=cfg["weight"], sparse=True).fetch(ch.name).tocsr()
        fea2 = trainUtils.buildmatrix(M, coords, w=cfg["w"], raw_M=raw_M, weights=ch.weights)
    assert np.array_equal(np.array(fea2), g[name + "/fea"])
    assert trainUtils.buildmatrix_from_pixels(ch.bin1, ch.bin2, ch.count, None, ch.n, coords[:5], w=cfg["w"]) is None


@pytest.mark.gpu
def test_depth_matches_dense_triu_sum(tmp_path, capsys):
    """`depth` (calculate_depth.py:25-28): device reduction = np.triu(raw, k).sum() for every
    upload path, and the sub-command prints the reference's three lines."""
    from peakachu_b200 import calculate_depth, cli, coolio
    from peakachu_b200 import rowpack
    case = Case("genome")
    path = case.write_cool(tmp_path)
    lib = coolio.PKCool(path)
    for ch in case.chroms[:2]:
        d = ch.bin2 - ch.bin1
        rp = np.searchsorted(ch.bin1, np.arange(ch.n + 1)).astype(np.int64)
        for k in (0, 3, 50):
            want = int(ch.count[d >= k].sum())
            assert calculate_depth.chromosome_depth((ch.bin1, ch.bin2, ch.count), ch.n, k) == want
            assert calculate_depth.chromosome_depth((rp, ch.bin2, ch.count), ch.n, k) == want
            assert calculate_depth.chromosome_depth(lib.upper_pixels_csr16(ch.name), ch.n, k) == want
            assert calculate_depth.chromosome_depth(lib.upper_pixels_rows(ch.name, 0), ch.n, k) == want
            assert calculate_depth.chromosome_depth(rowpack.pack_rows(rp, ch.bin2, ch.count, ch.n, 40), ch.n, k) == want
        perm = np.random.default_rng(1).permutation(ch.bin1.size)
        got = calculate_depth.chromosome_depth((ch.bin1[perm], ch.bin2[perm], ch.count[perm]), ch.n, 3)
        assert got == int(ch.count[d >= 3].sum())
    cli.run(["depth", "-p", path, "--min-dis", str(2 * case.cfg["res"])])
    out = capsys.readouterr().out.strip().splitlines()
    total = sum(int(c.count[(c.bin2 - c.bin1) >= 2].sum()) for c in case.chroms)
    assert out[-3] == "num of intra reads in your data: %d" % total
    assert out[-1].startswith("suggested model: ")
    assert out[:len(case.chroms)] == [c.name for c in case.chroms]


@pytest.mark.gpu
@pytest.mark.parametrize("name,min_dis", [("c5", 0), ("c5", 20000), ("genome", 0), ("gnames", 30000), ("tiny", 0), ("lowdepth", 50000)])
def test_depth_prints_what_the_reference_prints(name, min_dis, tmp_path, capsys):
    """`depth` against the reference's own stdout (tests/golden/depth.json, made by running the unmodified
    calculate_depth.main: make_depth_golden.py): chromosome names, contact total, human-equivalent depth and the
    suggested model, byte for byte."""
    import json

    from peakachu_b200 import cli
    from tests.cases import GOLDEN
    want = json.load(open(os.path.join(GOLDEN, "depth.json")))[name][str(min_dis)]
    path = Case(name).write_cool(tmp_path)
    capsys.readouterr()
    cli.run(["depth", "-p", path, "--min-dis", str(min_dis)])
    assert capsys.readouterr().out == want


@pytest.mark.gpu
def test_c5_as_specified_depth_selects_the_forest(tmp_path, capsys):
    """BASELINE.json configs[4] end to end: `depth` on the low-depth map suggests a model (calculate_depth.py:42-70),
    the forest is taken from a bank by that label, score_genome runs at --minimum-prob 0.6 and the reference's
    `pool` output at 0.6 and 0.9 (golden) consists of rows of our bedpe."""
    import json

    from peakachu_b200 import cli
    from tests.cases import GOLDEN
    case = Case("c5")
    cfg = case.cfg
    path = case.write_cool(tmp_path)
    capsys.readouterr()
    cli.run(["depth", "-p", path])
    printed = capsys.readouterr().out
    assert printed == json.load(open(os.path.join(GOLDEN, "depth.json")))["c5"]["0"]
    label = printed.strip().splitlines()[-1].split(": ", 1)[1]
    # the bank: the forest fitted at the map's own depth sits under the label the reference suggests for it,
    # forests of deeper maps under theirs (Case("genome") is a 300x map: "500 million")
    bank = {"10 million": case.pkl, "500 million": Case("genome").pkl, "450 million": Case("tiny").pkl}
    assert label == "10 million"
    out = os.path.join(str(tmp_path), "c5.bedpe")
    cli.run(["score_genome", "-p", path, "-m", bank[label], "-O", out, "-r", str(cfg["res"]), "-l", str(cfg["lower"]),
             "-u", str(cfg["upper"]), "--minimum-prob", str(cfg["min_prob"]), "--clr-weight-name", cfg["weight"],
             "-C"] + case.chroms_arg())
    txt = open(out).read()
    assert txt == case.bedpe
    rows = {tuple(ln.split("\t")[i] for i in (0, 1, 4)): ln.split("\t")[6:8] for ln in txt.splitlines()}
    for thr in (0.9, cfg["min_prob"]):
        pooled = [ln.split("\t") for ln in case.pool(thr).splitlines() if not ln.startswith("ERROR")]
        assert pooled or thr == 0.9
        for f in pooled:
            assert rows[(f[0], f[1], f[4])] == f[6:8]
    # a forest of the wrong depth gives a different loop set: the selection step matters
    out2 = os.path.join(str(tmp_path), "c5_wrong.bedpe")
    cli.run(["score_genome", "-p", path, "-m", bank["500 million"], "-O", out2, "-r", str(cfg["res"]), "-l", str(cfg["lower"]),
             "-u", str(cfg["upper"]), "--minimum-prob", str(cfg["min_prob"]), "--clr-weight-name", cfg["weight"],
             "-C"] + case.chroms_arg())
    assert open(out2).read() != txt


@pytest.mark.gpu
def test_full_size_c2_properties():
    """BASELINE configs[1] at full size (24,900 bins, 7.3 M band pixels, the 100-tree bench
    forest), where the oracle would take minutes: size-independent properties instead.
    (1) fused kernel == separate feature + forest kernels, record for record, bit for bit;
    (2) pruning (pixels that cannot exceed min_prob stop walking trees) changes nothing;
    (3) uint16 columns == int32 columns == unordered COO upload == packed pixel rows;
    (4) three band row tiles == the whole chromosome (the multi-GPU seam);
    (5) every record obeys prob > min_prob, lower <= y - x <= upper, and value == (w_x w_y) count;
    (6) a higher min_prob yields exactly the subset of the records above it."""
    from peakachu_b200 import _lib, shard, synth
    from peakachu_b200.forest import FlatForest
    from peakachu_b200.scoreUtils import Chromosome
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    flat = FlatForest.load(os.path.join(root, "bench_data", "c2_forest.npz"))
    ch = synth.make_chromosome("chr1", 24900, seed=1234, depth=300.0, band=330)
    n = ch.n
    rp = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)
    kw = dict(lower=6, upper=300, cname="chr1", res=10000, width=5)
    L = _lib.lib()
    X = Chromosome.from_csr(rp, ch.bin2, ch.count, ch.weights, n, flat, **kw)
    ref = X.score_records(0.5)
    assert ref[0].size > 10000
    try:
        _lib.check(L.pk_set_tuning(b"fused", 0))                              # (1)
        two = X.score_records(0.5)
        _lib.check(L.pk_set_tuning(b"fused", -1))
        _lib.check(L.pk_set_tuning(b"prune", 0))                              # (2)
        full = X.score_records(0.5)
    finally:
        _lib.check(L.pk_set_tuning(b"fused", -1))
        _lib.check(L.pk_set_tuning(b"prune", 1))
    for other in (two, full):
        assert all(np.array_equal(a, b) for a, b in zip(ref, other))
    hi = X.score_records(0.9)                                                 # (6)
    sel = ref[2] > 0.9
    assert all(np.array_equal(a[sel], b) for a, b in zip(ref, hi))
    X.close()
    N = Chromosome.from_csr16(rp, (ch.bin2 - ch.bin1).astype(np.uint16), ch.count.astype(np.uint16), ch.weights,
                              n, flat, **kw)                                  # (3)
    perm = np.random.default_rng(3).permutation(ch.bin1.size)
    U = Chromosome.from_pixels(ch.bin1[perm], ch.bin2[perm], ch.count[perm], ch.weights, n, flat,
                               sorted_pixels=False, **kw)
    from peakachu_b200 import rowpack
    P = Chromosome.from_rows(rowpack.pack_rows(rp, ch.bin2, ch.count, n, 320), ch.weights, n, flat, **kw)
    for Y in (N, U, P):
        got = Y.score_records(0.5)
        assert all(np.array_equal(a, b) for a, b in zip(ref, got))
        Y.close()

    class OneChrom:                                                           # (4)
        def upper_pixels_csr(self, k): return rp, ch.bin2, ch.count
        def weights(self, k, name): return ch.weights
        def nbins(self, k): return n
    tiles = shard.score_units(OneChrom(), [("chr1", 0, 9000), ("chr1", 9000, 17001), ("chr1", 17001, n)], flat,
                              correct="weight", lower=6, upper=300, res=10000, device=0, min_prob=0.5)
    tx, ty, tp, tv = shard.merge_tiles(sorted(tiles["chr1"], key=lambda q: q["row_begin"]))
    assert all(np.array_equal(a, b) for a, b in zip(ref, (tx, ty, tp, tv)))
    x, y, p, v = ref                                                          # (5)
    assert np.all(p > 0.5) and np.all(p <= 1.0)
    assert np.all(y - x >= 6) and np.all(y - x <= 300)
    assert np.all(np.lexsort((y, x)) == np.arange(x.size))
    from scipy import sparse
    C_ = sparse.csr_matrix((ch.count, (ch.bin1, ch.bin2)), shape=(n, n))
    cnt = np.asarray(C_[x, y]).ravel()
    assert np.array_equal(v, (ch.weights[x] * ch.weights[y]) * cnt)


@pytest.mark.parametrize("name", BIG_CASES)
def test_full_size_matches_reference(name, tmp_path):
    """BASELINE configs[1] at full size against the REFERENCE's own run on the same map and forest
    (tests/golden/c2.*, made by make_golden.py): expected curve bit for bit; candidate list, kept windows,
    float32 features (separate kernel and the tap inside the fused kernel), probabilities of every window
    and the bedpe text by checksum; the records column by column."""
    import hashlib
    from peakachu_b200 import score_chromosome
    from peakachu_b200.scoreUtils import DeviceForest

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    case = Case(name)
    cfg, ch = case.cfg, case.chroms[0]
    sh = case.meta["sha"][ch.name]
    X = _gpu_chromosome(case, ch, DeviceForest.of(case.forest, 0))
    assert np.array_equal(X.exp_arr, case.z[ch.name + "/exp_arr"])
    assert X.ridx.size == sh["n_candidates"]
    assert sha(X.ridx.astype(np.int32)) == sh["ridx"] and sha(X.cidx.astype(np.int32)) == sh["cidx"]
    keep, f32 = X.window_features()
    assert int(keep.sum()) == sh["n_windows"]
    clist = np.stack([X.ridx[keep], X.cidx[keep]], axis=1).astype(np.int32)
    assert sha(clist) == sh["clist"]
    assert sha(f32[keep]) == sh["fea32"]
    keep_f, f32_f = X.fused_window_features()
    assert np.array_equal(keep_f, keep) and sha(f32_f[keep_f]) == sh["fea32"]
    del f32, f32_f
    x, y, p, v = X.score_records(-1.0)                 # every window's probability, (x, y) order
    order = np.lexsort((clist[:, 1], clist[:, 0]))
    assert np.array_equal(x, clist[order, 0]) and np.array_equal(y, clist[order, 1])
    back = np.empty_like(p)
    back[order] = p                                    # back to the reference's window order
    assert sha(back) == sh["proba"]
    x, y, p, v = X.score_records(cfg["min_prob"])
    assert np.array_equal(x, case.z["records/x"]) and np.array_equal(y, case.z["records/y"])
    assert np.array_equal(p, case.z["records/prob"]) and np.array_equal(v, case.z["records/value"])
    X.close()
    out = os.path.join(str(tmp_path), "gpu.bedpe")
    score_chromosome.main(argparse.Namespace(path=case.write_cool(tmp_path), model=case.pkl, output=out,
                                             resolution=cfg["res"], lower=cfg["lower"], upper=cfg["upper"],
                                             minimum_prob=cfg["min_prob"], clr_weight_name=cfg["weight"],
                                             chrom=ch.name))
    txt = open(out).read()
    assert hashlib.sha256(txt.encode()).hexdigest() == case.meta["bedpe_sha"]


class _MemMap:
    """Chromosomes held in memory behind the interface shard.score_units reads a map through."""

    def __init__(self, nd_enc):
        self.nd_enc, self.ch = nd_enc, {}

    def add(self, ch):
        from peakachu_b200 import rowpack
        rp = np.searchsorted(ch.bin1, np.arange(ch.n + 1)).astype(np.int64)
        self.ch[ch.name] = dict(n=ch.n, w=ch.weights, rows=rowpack.pack_rows(rp, ch.bin2, ch.count, ch.n, self.nd_enc))

    def nbins(self, key): return self.ch[key]["n"]
    def weights(self, key, name): return self.ch[key]["w"]
    def upper_pixels_rows(self, key, nd_min): return self.ch[key]["rows"] if self.nd_enc >= nd_min else None


@pytest.mark.gpu
def test_full_size_genome_matches_reference():
    """BASELINE configs[2] at full size: score_genome's path (plan, engine, gather, bedpe text) on the hg19-shaped
    10 kb genome the bench scores -- 23 chromosomes, 303,641 bins, 88.5 M band pixels -- against the checksums of the
    REFERENCE's own bedpe for every chromosome (tests/golden/fullsize.json, made by make_fullsize_golden.py from the
    unmodified score_genome.main)."""
    import hashlib
    import json

    import bench
    from peakachu_b200 import shard, synth
    from peakachu_b200.forest import FlatForest
    from tests.cases import GOLDEN
    gold = json.load(open(os.path.join(GOLDEN, "fullsize.json")))["c3"]["chroms"]
    wl = bench.GENOMES["c3"]
    sizes = synth.hg19_bins(wl["res"])
    assert list(gold) == list(sizes)
    mm = _MemMap((wl["upper"] + 2 * wl["w"] + 1 + 31) // 32 * 32)
    for idx, (name, n) in enumerate(sizes.items()):
        ch = synth.make_chromosome(name, n, seed=5000 + idx, depth=wl["depth"], band=wl["band"])
        assert ch.checksum() == gold[name]["input_checksum"], name
        mm.add(ch)
    flat = FlatForest.load(os.path.join(bench.ROOT, "bench_data", wl["forest"] + "_forest.npz"))
    text = shard.score_chromosomes(mm, list(sizes), flat, correct="weight", lower=wl["lower"], upper=wl["upper"],
                                   res=wl["res"], min_prob=0.5, device=0)
    for name in sizes:
        assert text[name].count("\n") == gold[name]["rows"], name
        assert hashlib.sha256(text[name].encode()).hexdigest() == gold[name]["sha256"], name


@pytest.mark.gpu
def test_full_size_c4_chromosome_matches_reference(tmp_path):
    """BASELINE configs[3]'s shape at full size (49,850 bins at 5 kb, w = 7, -u 600, the 200-tree bench forest):
    the CLI's bedpe against the checksum of the reference's own (tests/golden/fullsize.json)."""
    import hashlib
    import json

    import bench
    from peakachu_b200 import coolio, score_chromosome
    from tests.cases import GOLDEN
    gold = json.load(open(os.path.join(GOLDEN, "fullsize.json")))["c4_chr1"]
    wl = bench.WORKLOADS["c4"]
    ch = bench.make_map(wl, seed=1234)
    assert ch.checksum() == gold["input_checksum"]
    path = os.path.join(str(tmp_path), "c4.pkcool")
    coolio.PKCool.write(path, [ch], wl["res"], rows_nd=(wl["upper"] + 2 * wl["w"] + 1 + 31) // 32 * 32)
    out = os.path.join(str(tmp_path), "c4.bedpe")
    score_chromosome.main(argparse.Namespace(path=path, model=os.path.join(bench.ROOT, "bench_data", wl["forest"] + ".pkl"),
                                             output=out, resolution=wl["res"], lower=wl["lower"], upper=wl["upper"],
                                             minimum_prob=0.5, clr_weight_name="weight", chrom=ch.name))
    txt = open(out).read()
    assert txt.count("\n") == gold["rows"]
    assert hashlib.sha256(txt.encode()).hexdigest() == gold["sha256"]


def test_batch_rule_whole_chromosome_and_row_tiles():
    """scoreUtils.py:104-108 on the fixture made for it (> 340,000 candidates; four reference batches keep
    2 / 1 / 0 / 3 windows): the whole-chromosome pass drops the lone window of the second batch on the
    device; with three row tiles the two windows of the first batch fall into different tiles, so each tile
    alone sees one -- only the per-batch sum over the tiles keeps them. Both give the reference's bedpe."""
    from peakachu_b200 import shard

    class OneChrom:
        def __init__(self, ch): self.ch = ch
        def upper_pixels(self, k): return self.ch.bin1, self.ch.bin2, self.ch.count
        def weights(self, k, name): return self.ch.weights
        def nbins(self, k): return self.ch.n

    case = Case("batchrule")
    cfg, ch = case.cfg, case.chroms[0]
    lib = OneChrom(ch)
    kw = dict(correct="weight", lower=cfg["lower"], upper=cfg["upper"], res=cfg["res"], device=0,
              min_prob=cfg["min_prob"])
    whole = shard.score_units(lib, [(ch.name, 0, ch.n)], case.forest, **kw)
    assert whole[ch.name][0]["batch_windows"].tolist() == [2, 1, 0, 3]
    assert whole[ch.name][0]["x"].size == 5                        # rule applied on the device
    edges = [0, 4666, 9333, ch.n]
    tiles = shard.score_units(lib, [(ch.name, a, b) for a, b in zip(edges[:-1], edges[1:])], case.forest, **kw)
    per_tile = [q["batch_windows"].tolist() for q in sorted(tiles[ch.name], key=lambda q: q["row_begin"])]
    assert per_tile == [[1, 0, 0, 1], [0, 1, 0, 2], [1, 0, 0, 0]]
    assert sum(q["x"].size for q in tiles[ch.name]) == 6           # a tile cannot apply the rule alone
    for res_ in (whole, tiles):
        assert shard.assemble_text([ch.name], [res_], cfg["res"])[ch.name] == case.bedpe


@pytest.mark.gpu
def test_full_size_c4_properties():
    """BASELINE configs[3] shape for one chromosome at full length (49,850 bins at 5 kb, upper=600,
    w=7, the 200-tree bench forest): fused kernel == separate kernels, pruning changes nothing,
    narrow columns == int32 columns, record invariants."""
    from peakachu_b200 import _lib, synth
    from peakachu_b200.forest import FlatForest
    from peakachu_b200.scoreUtils import Chromosome
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    flat = FlatForest.load(os.path.join(root, "bench_data", "c4_forest.npz"))
    ch = synth.make_chromosome("chr1", 49850, seed=1234, depth=300.0, band=640)
    n = ch.n
    rp = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)
    kw = dict(lower=6, upper=600, cname="chr1", res=5000, width=7)
    L = _lib.lib()
    X = Chromosome.from_csr(rp, ch.bin2, ch.count, ch.weights, n, flat, **kw)
    assert X.lower == 8                                     # max(lower, w + 1), scoreUtils.py:13
    ref = X.score_records(0.5)
    assert ref[0].size > 10000
    try:
        _lib.check(L.pk_set_tuning(b"fused", 0))
        two = X.score_records(0.5)
        _lib.check(L.pk_set_tuning(b"fused", -1))
        _lib.check(L.pk_set_tuning(b"prune", 0))
        full = X.score_records(0.5)
    finally:
        _lib.check(L.pk_set_tuning(b"fused", -1))
        _lib.check(L.pk_set_tuning(b"prune", 1))
    for other in (two, full):
        assert all(np.array_equal(a, b) for a, b in zip(ref, other))
    X.close()
    N = Chromosome.from_csr16(rp, (ch.bin2 - ch.bin1).astype(np.uint16), ch.count.astype(np.uint16), ch.weights,
                              n, flat, **kw)
    got = N.score_records(0.5)
    assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    N.close()
    x, y, p, v = ref
    assert np.all(p > 0.5) and np.all(y - x >= 8) and np.all(y - x <= 600)
    assert np.all(np.lexsort((y, x)) == np.arange(x.size))


@pytest.mark.gpu
def test_long_chromosome_matches_oracle(tmp_path):
    """70,000 bins (longer than 57 * 1024: the per-diagonal sums take their large-table variant,
    numpy's pairwise tree is two levels deeper): expected curve, candidates and records bit-exact
    against the oracle."""
    import io
    from contextlib import redirect_stdout
    from oracle import peakachu_oracle as po
    from peakachu_b200 import coolio, synth
    from peakachu_b200.scoreUtils import Chromosome
    case = Case("tiny")
    ch = synth.make_chromosome("chr1", 70000, seed=9, depth=40.0, band=90, n_loops=200, loop_max=60)
    X = Chromosome.from_pixels(ch.bin1, ch.bin2, ch.count, ch.weights, ch.n, case.forest, lower=6, upper=60,
                               cname="chr1", res=10000, width=5, sorted_pixels=True)
    x, y, p, v = X.score_records(0.5)
    path = os.path.join(str(tmp_path), "long.pkcool")
    coolio.PKCool.write(path, [ch], 10000)
    lib = coolio.Cooler(path)
    M = po.tocsr(lib.matrix(balance="weight", sparse=True).fetch("chr1"))
    raw = po.tocsr(lib.matrix(balance=False, sparse=True).fetch("chr1"))
    O = po.Chromosome(M, model=case.model(), raw_M=raw, weights=ch.weights, lower=6, upper=60, cname="chr1",
                      res=10000, width=5)
    assert np.array_equal(X.exp_arr, O.exp_arr)
    assert np.array_equal(X.ridx, O.ridx) and np.array_equal(X.cidx, O.cidx)
    with redirect_stdout(io.StringIO()):
        prob, val = O.score(0.5)
    r, c = prob.nonzero()
    assert np.array_equal(x, r) and np.array_equal(y, c)
    assert np.array_equal(p, np.asarray(prob[r, c]).ravel()) and np.array_equal(v, np.asarray(val[r, c]).ravel())
    X.close()


@pytest.mark.gpu
@pytest.mark.parametrize("raw", [False, True])
def test_rows_without_the_far_pixels_valid_does_not_need(raw, tmp_path, monkeypatch):
    """From a cooler file the scoring path uploads packed rows whose far lists hold only the pixels that make a bin
    valid which no pixel inside the band does (pk_rows_pack far_mode 1). On a map with pixels far beyond the band,
    NaN / overflowing weights and bins that are valid through far pixels alone, the `valid`-dependent expected
    curve, the candidates and the records equal those of the rows with every far pixel, and of the plain columns."""
    import copy
    from peakachu_b200 import coolio, rowpack
    from peakachu_b200.scoreUtils import Chromosome
    from tests import h5write
    case = Case("tiny")
    ch = copy.copy(case.chroms[0])
    rng = np.random.default_rng(5)
    n = ch.n
    # bins 100..119 lose every pixel within 80 bins (rows and columns) and get pixels at distance >= 90 instead
    lonely = np.arange(100, 120)
    keep = ~((np.isin(ch.bin1, lonely) | np.isin(ch.bin2, lonely)) & (ch.bin2 - ch.bin1 < 80))
    fb1 = np.repeat(lonely, 3)
    fb2 = fb1 + rng.integers(90, 200, fb1.size)
    far_b1 = rng.integers(0, n - 150, 400)
    far_b2 = far_b1 + rng.integers(75, 150, 400)
    b1 = np.concatenate([ch.bin1[keep], fb1, far_b1]); b2 = np.concatenate([ch.bin2[keep], fb2, far_b2])
    cnt = np.concatenate([ch.count[keep], np.full(fb1.size, 2), rng.integers(1, 4, 400)])
    key, first = np.unique(b1.astype(np.int64) * n + b2, return_index=True)
    ch.bin1, ch.bin2, ch.count = (key // n).astype(np.int32), (key % n).astype(np.int32), cnt[first].astype(np.int32)
    w = ch.weights.copy()
    w[105] = np.nan; w[110] = 1e200; w[300] = 1e200; w[111] = np.inf
    ch.weights = w
    path = str(tmp_path / "far.cool")
    h5write.write_cool(path, [ch], 10000, chunk=3000)
    lib = coolio.open_map(path)
    weights = None if raw else lib.weights(ch.name, "weight")
    kw = dict(lower=6, upper=60, cname="chr1", res=10000, width=5)
    got = {}
    for mode in ("1", "0", "csr32"):
        monkeypatch.setenv("PEAKACHU_B200_SLIM_FAR", "0" if mode == "csr32" else mode)
        X = Chromosome.from_map(lib, ch.name, weights, case.forest, encoding="csr32" if mode == "csr32" else None, **kw)
        got[mode] = (X.exp_arr.copy(), X.ridx.copy(), X.cidx.copy()) + tuple(X.score_records(0.5))
        X.close()
    nd = 60 + 2 * 5 + 1
    slim, full = lib.upper_pixels_rows(ch.name, nd, scoring_weights=weights), lib.upper_pixels_rows(ch.name, nd)
    assert 0 < rowpack.header(slim)["n_far"] < rowpack.header(full)["n_far"]
    assert got["1"][1].size > 0 and got["1"][3].size > 0
    for mode in ("0", "csr32"):
        for a, b in zip(got["1"], got[mode]):
            assert np.array_equal(a, b, equal_nan=True), mode
