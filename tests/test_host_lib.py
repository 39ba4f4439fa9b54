"""CPU-side checks of the C-ABI library: it loads, exports every symbol the
header declares, and its host-only arithmetic (Poisson decision table, expected
curve fit) agrees with scipy / scikit-learn. No CUDA device needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from scipy import stats
from sklearn.isotonic import IsotonicRegression

from peakachu_b200 import _lib, shard
from peakachu_b200.scoreUtils import format_bedpe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "peakachu_b200.h")).read()
    declared = set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "libpeakachu_b200.so lacks %s" % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert _lib.lib().pk_abi_version() == 1


def test_poisson_table_decides_like_scipy():
    L = _lib.lib()
    kmax = 3000
    crit = np.zeros(kmax + 1)
    _lib.check(L.pk_poisson_critical_mu(kmax, _lib.ptr(crit, _lib.c_f64p)))
    ks = np.arange(kmax + 1)
    assert np.all(np.diff(crit) > 0)
    # just below / above the critical mean scipy agrees on the side of 0.01
    assert np.all(stats.poisson(crit * (1 - 1e-11)).sf(ks) < 0.01)
    assert np.all(stats.poisson(crit * (1 + 1e-11)).sf(ks) >= 0.01)
    # random (k, mu): identical decisions
    rng = np.random.default_rng(3)
    k = rng.integers(1, kmax, 200000)
    mu = crit[k] * np.exp(rng.normal(0, 0.05, k.size))
    assert np.array_equal(mu < crit[k], stats.poisson(mu).sf(k) < 0.01)


def test_expected_fit_matches_sklearn_bitwise():
    L = _lib.lib()
    rng = np.random.default_rng(4)
    for trial in range(300):
        n = int(rng.integers(3, 700))
        mean = 300.0 / (1.0 + np.arange(n)) ** rng.uniform(0.5, 1.5) * np.exp(rng.uniform(0, 0.6) * rng.standard_normal(n))
        if trial % 3 == 0:
            mean = np.round(mean, 1)
        cnt = rng.integers(5, 3000, n).astype(np.int64)
        if trial % 4 == 0:
            mean[rng.random(n) < 0.2] = 0.0
        s = mean * cnt                                   # library divides sum by count
        e = np.where(cnt > 10, s / cnt, 0.0)
        out = np.zeros(n)
        rc = L.pk_fit_expected(_lib.ptr(s, _lib.c_f64p), _lib.ptr(cnt, _lib.c_i64p), n, _lib.ptr(out, _lib.c_f64p))
        d = np.where(e > 0)[0]
        if d.size == 0:
            assert rc != 0
            continue
        assert rc == 0
        IR = IsotonicRegression(increasing=False, out_of_bounds="clip").fit(d, e[d])
        assert np.array_equal(out, IR.predict(list(range(n)))), trial


def test_bedpe_float_text_is_numpy_str():
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.random(2000), rng.random(2000) * 1e-3, rng.random(500) * 1e-6,
                           rng.gamma(2, 50, 500), [0.5, 1.0, 0.1 + 0.2, 1e-5, 123456789.125, 1e16, 1e-300]])
    txt = format_bedpe("chr1", 10000, np.arange(vals.size, dtype=np.int32), np.arange(vals.size, dtype=np.int32) + 7,
                       vals, vals[::-1])
    rows = txt.splitlines()
    for i in (0, 17, 4999, vals.size - 1, vals.size - 2, vals.size - 3):
        r, c = np.int32(i), np.int32(i + 7)
        want = "\t".join(map(str, ["chr1", r * 10000, (r + 1) * 10000, "chr1", c * 10000, (c + 1) * 10000,
                                     np.float64(vals[i]), np.float64(vals[::-1][i])]))
        assert rows[i] == want
    assert all(str(np.float64(v)) == repr(float(v)) for v in vals)
    # every row, plus awkward magnitudes, against Python/numpy formatting
    extra = np.array([1e15, 1e16, 9.999999999999999e15, 1e-4, 9.999e-5, 1.5e-7, 5e-324, 1.7976931348623157e308,
                      123456.0, 0.0, 1.0, 100.0, 2.5e-5, 0.30000000000000004])
    rng2 = np.random.default_rng(6)
    wild = np.exp(rng2.uniform(-60, 60, 20000)) * rng2.choice([1.0, 0.5, 3.0], 20000)
    allv = np.concatenate([vals, extra, wild])
    txt = format_bedpe("chrX", 5000, np.arange(allv.size, dtype=np.int32), np.arange(allv.size, dtype=np.int32),
                       allv, allv)
    for i, row in enumerate(txt.splitlines()):
        f = row.split("\t")
        assert f[6] == str(np.float64(allv[i])) == f[7], (i, allv[i], f[6])
        assert f[1] == str(np.int32(i) * 5000) and f[2] == str((np.int32(i) + 1) * 5000)


def test_plan_covers_every_row_once_and_balances():
    sizes = {"chr%d" % i: n for i, n in enumerate([24926, 24320, 19803, 15900, 9036, 6303, 4813, 700], 1)}
    for world in (1, 2, 4, 8):
        asg = shard.plan(sizes, world, 6, 300, 5)
        seen = {}
        for units in asg:
            for k, a, b in units:
                seen.setdefault(k, []).append((a, b))
        assert set(seen) == set(sizes)
        for k, tiles in seen.items():
            tiles.sort()
            assert tiles[0][0] == 0 and tiles[-1][1] == sizes[k]
            assert all(t0[1] == t1[0] for t0, t1 in zip(tiles, tiles[1:]))
        loads = [sum(shard.band_pixels(sizes[k], 6, 300, 5) * (b - a) / sizes[k] for k, a, b in u) for u in asg]
        assert max(loads) <= 1.35 * (sum(loads) / world) + 1


def test_expected_fit_mode_follows_the_environment(monkeypatch):
    """The device fit is pinned to the arithmetic of one scikit-learn / scipy / numpy generation: with other
    versions installed `auto` warns (once) and still fits on the device; PEAKACHU_B200_EXPECTED selects the mode."""
    import warnings

    from peakachu_b200 import _lib
    have = _lib.installed_versions()
    assert set(have) == set(_lib.PINNED_VERSIONS)
    monkeypatch.delenv("PEAKACHU_B200_EXPECTED", raising=False)
    monkeypatch.setattr(_lib, "_expected_mode", None)
    monkeypatch.setattr(_lib, "PINNED_VERSIONS", dict(have))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert _lib.expected_mode() == "device"                       # matching environment: silent
    monkeypatch.setattr(_lib, "_expected_mode", None)
    monkeypatch.setattr(_lib, "PINNED_VERSIONS", dict(have, sklearn="0.0"))
    with pytest.warns(RuntimeWarning, match="PEAKACHU_B200_EXPECTED=host"):
        assert _lib.expected_mode() == "device"
    for mode in ("host", "device"):
        monkeypatch.setattr(_lib, "_expected_mode", None)
        monkeypatch.setenv("PEAKACHU_B200_EXPECTED", mode)
        assert _lib.expected_mode() == mode
    monkeypatch.setattr(_lib, "_expected_mode", None)
    monkeypatch.setenv("PEAKACHU_B200_EXPECTED", "sometimes")
    with pytest.raises(ValueError):
        _lib.expected_mode()
    monkeypatch.setattr(_lib, "_expected_mode", None)


def test_dropin_constructor_refuses_a_differently_balanced_matrix(tmp_path):
    """Chromosome(M, model, raw_M, weights): the device balances raw_M with `weights` itself, so an M that is not
    (w[r] * w[c]) * raw -- another weight column, a scaled matrix -- is refused on the host, before any device work."""
    from scipy import sparse

    from peakachu_b200 import synth
    from peakachu_b200.scoreUtils import Chromosome, _check_balanced_matrix
    ch = synth.make_chromosome("chr1", 300, seed=3, depth=100.0, band=60, n_loops=10, loop_max=40)
    raw = sparse.coo_matrix((ch.count, (ch.bin1, ch.bin2)), shape=(ch.n, ch.n)).tocsr()
    raw = raw + sparse.triu(raw, k=1).T
    w = ch.weights
    rc = raw.tocoo()
    good = sparse.csr_matrix(((w[rc.row] * w[rc.col]) * rc.data, (rc.row, rc.col)), shape=raw.shape)
    b1, b2 = ch.bin1.astype(np.int32), ch.bin2.astype(np.int32)
    _check_balanced_matrix(good, b1, b2, ch.count, w)                       # the product itself: accepted
    with pytest.raises(ValueError, match="balances raw_M"):
        _check_balanced_matrix(good * 1.5, b1, b2, ch.count, w)
    with pytest.raises(ValueError, match="balances raw_M"):
        Chromosome(good, model=None, raw_M=raw, weights=w * 1.01)           # raised before the device is touched


def test_score_chromosome_reports_the_model_before_the_map(tmp_path):
    """score_chromosome.main unpickles the model on a thread while the map is read; like the reference
    (score_chromosome.py:14 comes before :33) a missing model is the error of a run that lacks both."""
    import argparse

    from peakachu_b200 import score_chromosome
    ns = argparse.Namespace(path=str(tmp_path / "none.cool"), model=str(tmp_path / "none.pkl"), output=str(tmp_path / "o.bedpe"),
                            chrom="chr1", lower=6, upper=300, minimum_prob=0.5, resolution=10000, clr_weight_name="weight")
    with pytest.raises(FileNotFoundError, match="none.pkl"):
        score_chromosome.main(ns)
    import os
    ns.model = os.path.join(os.path.dirname(__file__), "golden", "tiny_forest.npz")
    with pytest.raises(FileNotFoundError, match="none.cool"):
        score_chromosome.main(ns)
