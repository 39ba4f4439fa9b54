"""Real `.cool` ingestion: the HDF5 subset reader and the cooler
adapter, on CPU. The reference's entry is `cooler.Cooler(uri)` + `matrix(...).fetch(chrom)` +
`bins().fetch(chrom)[weight]` (score_chromosome.py:33-44)."""
import glob
import os

import numpy as np
import pytest

from peakachu_b200 import coolio, h5mini, synth
from tests import h5write


def _genome(seed=3):
    return synth.make_genome({"chr1": 900, "chr2": 640, "chrX": 410}, seed=seed, depth=60.0)


def _trans(chroms, rng, k=500):
    nb = np.array([c.n for c in chroms])
    off = np.concatenate([[0], np.cumsum(nb)])
    out = []
    for _ in range(k):
        i = rng.integers(0, len(chroms) - 1)
        j = rng.integers(i + 1, len(chroms))
        out.append((off[i] + rng.integers(0, nb[i]), off[j] + rng.integers(0, nb[j]), rng.integers(1, 5)))
    return np.unique(np.array(out, dtype=np.int64), axis=0)


def test_reader_on_a_libhdf5_written_file():
    """The one HDF5 file in the image that libhdf5 itself wrote (MATLAB 7.3, 512-byte user block,
    symbol-table root group, version-1 object header, fixed-string attribute)."""
    import scipy.io
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat"))
    if not hits:
        pytest.skip("scipy test data not installed")
    with h5mini.File(hits[0]) as f:
        assert f.root.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
        np.testing.assert_array_equal(d.read().ravel(), np.arange(9) * (np.pi / 4))     # scipy's test_mio: 0 .. 2 pi
        assert d.attrs["MATLAB_class"] == b"double"


@pytest.mark.parametrize("latest", [False, True])
@pytest.mark.parametrize("userblock,group,chunk", [(0, "", 4096), (512, "resolutions/10000", 1000), (0, "", 37)])
def test_cool_round_trip(tmp_path, userblock, group, chunk, latest):
    """`latest`: the same cooler in HDF5's newer on-disk structures (superblock 2, OHDR object headers,
    link-message groups, version-4 layouts with fixed-array / single-chunk indexes; h5py libver='latest')."""
    chroms = _genome()
    rng = np.random.default_rng(0)
    path = str(tmp_path / "t.cool")
    ref = h5write.write_cool(path, chroms, 10000, trans=_trans(chroms, rng), group=group, chunk=chunk, latest=latest,
                             userblock=userblock, extra_bins={"KR": np.arange(sum(c.n for c in chroms), dtype=np.float64)})
    uri = path + ("::/" + group if group else "")
    lib = coolio.open_map(uri)
    assert isinstance(lib, coolio.H5Cool)
    # (latest, chunk 37: more than 1024 chunks per pixel column, so the fixed-array index is paged like libhdf5 pages it)
    assert lib.chromnames == [c.name for c in chroms]
    assert lib.binsize == 10000
    off = 0
    for c in chroms:
        assert lib.nbins(c.name) == c.n
        b1, b2, cnt = lib.upper_pixels(c.name)
        order = np.lexsort((c.bin2, c.bin1))
        np.testing.assert_array_equal(b1, c.bin1[order])
        np.testing.assert_array_equal(b2, c.bin2[order])
        np.testing.assert_array_equal(cnt, c.count[order])
        assert b1.dtype == np.int32 and cnt.dtype == np.int32
        rp, b2c, cntc = lib.upper_pixels_csr(c.name)
        assert rp.dtype == np.int64 and rp.size == c.n + 1 and rp[0] == 0 and rp[-1] == b2c.size
        np.testing.assert_array_equal(np.repeat(np.arange(c.n), np.diff(rp)), b1)
        narrow = lib.upper_pixels_csr16(c.name)
        assert narrow is not None
        np.testing.assert_array_equal(narrow[1].astype(np.int64) + b1, b2)
        np.testing.assert_array_equal(narrow[2], cnt)
        assert lib.weights_divisive("KR") and not lib.weights_divisive("weight")      # cooler's column attribute
        w = lib.weights(c.name, "weight")
        np.testing.assert_array_equal(w, c.weights)           # NaN positions included
        np.testing.assert_array_equal(lib.weights(c.name, "KR"), np.arange(off, off + c.n, dtype=np.float64))
        off += c.n
    with pytest.raises(KeyError):
        lib.weights("chr1", "VC")
    with pytest.raises(KeyError):
        lib.upper_pixels("chr9")
    lib.close()
    # the raw columns, chunk by chunk and in partial reads
    with h5mini.File(path) as f:
        g = f[group] if group else f.root
        assert sorted(g.keys()) == ["bins", "chroms", "indexes", "pixels"]
        assert g.attrs["nnz"] == ref["bin1"].size and g.attrs["format"] == b"HDF5::Cooler"
        d = g["pixels/bin1_id"]
        np.testing.assert_array_equal(d.read(), ref["bin1"])
        for lo, hi in [(0, 1), (chunk - 1, chunk + 1), (5, 3 * chunk + 7), (ref["bin1"].size - 3, ref["bin1"].size + 10)]:
            np.testing.assert_array_equal(d.read(lo, hi), ref["bin1"][lo:hi])
            np.testing.assert_array_equal(d[lo:hi], ref["bin1"][lo:hi])
        np.testing.assert_array_equal(g["pixels/count"].read(), ref["count"])     # fletcher32 + shuffle + deflate
        np.testing.assert_array_equal(g["bins/weight"].read(), ref["weights"])   # one chunk with deflate masked out
        assert g["bins/weight"].attrs["ignore_diags"] == 2
        np.testing.assert_array_equal(g["indexes/chrom_offset"].read(), ref["chrom_offset"])  # contiguous
        np.testing.assert_array_equal(g["bins/chrom"].read(), np.repeat(np.arange(3), [c.n for c in chroms]))  # enum


def test_same_pixels_as_the_pkcool_container(tmp_path):
    """A .cool and a .pkcool of the same map hand the scoring path identical columns."""
    chroms = _genome(seed=5)
    p1, p2 = str(tmp_path / "a.cool"), str(tmp_path / "a.pkcool")
    h5write.write_cool(p1, chroms, 10000)
    coolio.PKCool.write(p2, chroms, 10000)
    a, b = coolio.open_map(p1), coolio.open_map(p2)
    assert type(a) is coolio.H5Cool and type(b) is coolio.PKCool
    for c in chroms:
        for fa, fb in [(a.upper_pixels, b.upper_pixels), (a.upper_pixels_csr, b.upper_pixels_csr), (a.upper_pixels_csr16, b.upper_pixels_csr16)]:
            for x, y in zip(fa(c.name), fb(c.name)):
                np.testing.assert_array_equal(x, y)
                assert x.dtype == y.dtype
        np.testing.assert_array_equal(a.weights(c.name, "weight"), b.weights(c.name, "weight"))


@pytest.mark.parametrize("latest", [False, True])
def test_layout_and_type_variants(tmp_path, latest):
    """Compact and contiguous layouts, big-endian and narrow integer columns, float32, empty datasets,
    groups wider than one symbol-table node, multi-level chunk B-trees; in the newer format: implicit
    (unfiltered), single-chunk and fixed-array chunk indexes."""
    W = (h5write.Writer2 if latest else h5write.Writer)()
    rng = np.random.default_rng(1)
    big = rng.integers(-2**40, 2**40, 70000)
    kids = {
        "compact": W.dataset(np.arange(7, dtype=np.int16), compact=True),
        "be": W.dataset(np.arange(11, dtype=">i4")),
        "f4": W.dataset(np.linspace(0, 1, 33, dtype=np.float32), chunk=8, gzip=1),
        "u1": W.dataset(np.arange(200, dtype=np.uint8), chunk=64, shuffle=True),
        "empty": W.dataset(np.zeros(0, dtype=np.int64), chunk=16, gzip=6),
        "deep": W.dataset(big, chunk=100, gzip=1, shuffle=True),          # 700 chunks: two B-tree levels / a fixed array
        "plain": W.dataset(np.arange(1000, dtype=np.int32), chunk=128),    # chunked, no filter (implicit index when latest)
        "one": W.dataset(np.arange(50, dtype=np.int64), chunk=64, gzip=4, shuffle=True),   # a single, partial chunk
        "one_plain": W.dataset(np.arange(50, dtype=np.int64), chunk=64),
    }
    for i in range(20):                                                    # > 8 links: several SNODs
        kids["col%02d" % i] = W.dataset(np.full(3, i, dtype=np.int32))
    root = W.group(kids, attrs={"k": np.float64(2.5)})
    path = str(tmp_path / "v.h5")
    W.finish(root, path)
    with h5mini.File(path) as f:
        assert len(f.root.keys()) == 29 and f.root.attrs["k"] == 2.5
        np.testing.assert_array_equal(f["plain"].read(), np.arange(1000))
        np.testing.assert_array_equal(f["plain"].read(120, 300), np.arange(120, 300))
        np.testing.assert_array_equal(f["one"].read(), np.arange(50))
        np.testing.assert_array_equal(f["one_plain"][3:40], np.arange(3, 40))
        if latest:
            assert [f[k]._layout[0] for k in ("deep", "plain", "one", "one_plain")] == ["farray", "implicit", "single", "single"]
        np.testing.assert_array_equal(f["compact"].read(), np.arange(7))
        assert f["be"].read().dtype == np.dtype("=i4")
        np.testing.assert_array_equal(f["be"].read(), np.arange(11))
        np.testing.assert_array_equal(f["f4"].read(), np.linspace(0, 1, 33, dtype=np.float32))
        np.testing.assert_array_equal(f["u1"][10:150], np.arange(10, 150))
        assert f["empty"].read().size == 0
        np.testing.assert_array_equal(f["deep"].read(), big)
        np.testing.assert_array_equal(f["deep"].read(12345, 54321), big[12345:54321])
        for i in range(20):
            np.testing.assert_array_equal(f["col%02d" % i].read(), np.full(3, i))
        with pytest.raises(KeyError):
            f["nope"]


def test_not_hdf5(tmp_path):
    p = tmp_path / "x.cool"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(RuntimeError):
        coolio.open_map(str(p))
    with pytest.raises(FileNotFoundError):
        coolio.open_map(str(tmp_path / "missing.cool"))


def _mini_cool(path, b1, b2, cnt, n=50):
    """A one-chromosome cooler built column by column (lets a test plant odd columns)."""
    W = h5write.Writer()
    off = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    z = dict(chunk=64, gzip=6, shuffle=True)
    g = W.group({
        "chroms": W.group({"name": W.dataset(np.array([b"chrZ"])), "length": W.dataset(np.array([n * 1000], np.int32))}),
        "bins": W.group({"start": W.dataset(np.arange(n, dtype=np.int32) * 1000, **z),
                         "end": W.dataset(np.arange(1, n + 1, dtype=np.int32) * 1000, **z),
                         "weight": W.dataset(np.ones(n), **z)}),
        "pixels": W.group({"bin1_id": W.dataset(np.asarray(b1, np.int64), **z), "bin2_id": W.dataset(np.asarray(b2, np.int64), **z),
                           "count": W.dataset(np.asarray(cnt), **z)}),
        "indexes": W.group({"chrom_offset": W.dataset(np.array([0, n], np.int64)), "bin1_offset": W.dataset(off, **z)})})
    W.finish(g, path)


def test_odd_pixel_columns_fail_loudly(tmp_path):
    """Float counts are accepted when they are whole numbers (some coolers store them so); fractional
    counts (a balanced or normalised matrix in place of raw counts) and pixels below the diagonal are
    refused: the Poisson filter of scoreUtils.py:59-60 needs raw upper-triangle counts."""
    b1 = np.repeat(np.arange(40), 3)
    b2 = b1 + np.tile(np.arange(3), 40)
    p = str(tmp_path / "f.cool")
    _mini_cool(p, b1, b2, np.arange(1, 121, dtype=np.float64))
    lib = coolio.open_map(p)
    assert lib.binsize == 1000                       # no bin-size attribute: taken from bins/start, bins/end
    r1, r2, cnt = lib.upper_pixels("chrZ")
    assert cnt.dtype == np.int32 and np.array_equal(cnt, np.arange(1, 121)) and np.array_equal(r2, b2)
    _mini_cool(p, b1, b2, np.arange(1, 121) + 0.5)
    with pytest.raises(ValueError, match="non-integer"):
        coolio.open_map(p).upper_pixels("chrZ")
    b2l = b2.copy()
    b2l[30] = b1[30] - 2
    _mini_cool(p, b1, b2l, np.ones(120, np.int32))
    with pytest.raises(ValueError, match="below the diagonal"):
        coolio.open_map(p).upper_pixels("chrZ")
    _mini_cool(p, b1, b2, np.full(120, 2**31 + 5, np.int64))
    with pytest.raises(ValueError, match="outside int32"):
        coolio.open_map(p).upper_pixels("chrZ")


@pytest.mark.parametrize("latest", [False, True])
def test_native_chunk_decoder_equals_the_python_one(tmp_path, latest, monkeypatch):
    """Large reads of 1-D columns are decoded by the library (pk_h5_decode_chunks: zlib + un-shuffle on host threads,
    straight from the mapped file); small reads and other pipelines by the Python loop. Same arrays either way, for
    every filter pipeline the native path takes, whole columns and ranges that start and end inside chunks."""
    W = (h5write.Writer2 if latest else h5write.Writer)()
    rng = np.random.default_rng(7)
    cols = {"i8": rng.integers(-2**60, 2**60, 400000),                                   # incompressible: > 1 MB of chunks
            "i4": rng.integers(-2**31, 2**31 - 1, 700000).astype(np.int32),
            "f8": rng.normal(size=300000),
            "u2": rng.integers(0, 65535, 1500000).astype(np.uint16)}
    kids = {"i8": W.dataset(cols["i8"], chunk=8192, gzip=6, shuffle=True),
            "i4": W.dataset(cols["i4"], chunk=50000, gzip=1),
            "f8": W.dataset(cols["f8"], chunk=4096, gzip=4, shuffle=True, fletcher=True),
            "u2": W.dataset(cols["u2"], chunk=100000, gzip=6, fletcher=True)}
    path = str(tmp_path / "big.h5")
    W.finish(W.group(kids), path)
    calls = []
    real = h5mini.Dataset._decode_native

    def spy(self, hits, lo, hi, out):
        ok = real(self, hits, lo, hi, out)
        calls.append(ok)
        return ok
    with h5mini.File(path) as f:
        for name, want in cols.items():
            for lo, hi in [(0, want.size), (12345, want.size - 777), (want.size // 2, want.size // 2 + 5)]:
                monkeypatch.setattr(h5mini.Dataset, "_decode_native", spy)
                got = f[name].read(lo, hi)
                monkeypatch.setattr(h5mini.Dataset, "_decode_native", lambda self, *a: False)
                ref = f[name].read(lo, hi)
                assert got.dtype == ref.dtype and np.array_equal(got, ref) and np.array_equal(got, want[lo:hi]), (name, lo, hi)
    assert calls.count(True) == 8 and calls.count(False) == 4          # whole columns and long ranges natively, 5-element reads not


def test_lookup3_known_answers():
    """HDF5's metadata checksum (Jenkins lookup3 hashlittle) against the answers printed in lookup3.c's self-test."""
    assert h5mini.lookup3(b"") == 0xdeadbeef
    assert h5mini.lookup3(b"", 0xdeadbeef) == 0xbd5b7dde
    assert h5mini.lookup3(b"Four score and seven years ago") == 0x17770551
    assert h5mini.lookup3(b"Four score and seven years ago", 1) == 0xcd628161


def test_paged_fixed_array_chunk_index(tmp_path):
    """Version-4 layout with a fixed-array index of more entries than a page holds: pages behind the data block,
    a bitmap of the pages that exist, a checksum per page. Uninitialised pages read as zeros (absent chunks);
    a page whose checksum does not match fails loudly."""
    rng = np.random.default_rng(3)
    a = rng.integers(1, 1000, 1003).astype(np.int32)
    for bits, absent in ((2, ()), (3, (1,)), (4, (0, 6)), (10, ())):
        W = h5write.Writer2()
        path = str(tmp_path / ("p%d.h5" % bits))
        W.finish(W.group({"x": W.dataset(a, chunk=10, gzip=4, shuffle=True, page_bits=bits, absent_pages=absent)}), path)
        want = a.copy()
        for pg in absent:
            want[pg * (1 << bits) * 10:(pg + 1) * (1 << bits) * 10] = 0
        with h5mini.File(path) as f:
            np.testing.assert_array_equal(f["x"].read(), want)
            np.testing.assert_array_equal(f["x"].read(95, 407), want[95:407])
    # flip one byte inside the second page of the 2-bit file: the reader must not hand out chunk addresses from it
    path = str(tmp_path / "p2.h5")
    raw = bytearray(open(path, "rb").read())
    at = raw.index(b"FADB")
    npages = -(-101 // 4)
    first_page = at + 6 + 8 + (npages + 7) // 8 + 4
    raw[first_page + (4 * 16 + 4) + 9] ^= 0x40
    bad = str(tmp_path / "bad.h5")
    open(bad, "wb").write(raw)
    with h5mini.File(bad) as f:
        with pytest.raises(h5mini.H5Error, match="page 1 checksum"):
            f["x"].read()


@pytest.mark.parametrize("latest", [False, True])
def test_packed_rows_straight_from_the_file(tmp_path, latest):
    """H5Cool.upper_pixels_rows: the file's columns (genome-wide int64 bin2 ids, trans pixels, the file's count type)
    packed by pk_rows_pack are the blob rowpack.pack_rows makes of the chromosome's cis columns; the plain columns
    asked for afterwards come from the same decoded arrays; odd files fail as loudly on this path."""
    from peakachu_b200 import rowpack
    chroms = _genome()
    rng = np.random.default_rng(1)
    path = str(tmp_path / "t.cool")
    h5write.write_cool(path, chroms, 10000, trans=_trans(chroms, rng), chunk=500, latest=latest)
    lib = coolio.open_map(path)
    for c in chroms:
        for nd in (1, 40, 331):
            blob = lib.upper_pixels_rows(c.name, nd)
            order = np.lexsort((c.bin2, c.bin1))
            rp = np.searchsorted(c.bin1[order], np.arange(c.n + 1)).astype(np.int64)
            assert np.array_equal(blob, rowpack.pack_rows(rp, c.bin2[order], c.count[order], c.n, nd))
        rp2, b2, cnt = lib.upper_pixels_csr(c.name)          # served from the columns decoded above
        assert np.array_equal(rp2, rp) and np.array_equal(b2, c.bin2[order]) and np.array_equal(cnt, c.count[order])
    # too many pixels to hold whole: the caller is sent to the block-wise columns
    lib2 = coolio.open_map(path)
    lib2.RAW_LIMIT = 10
    assert lib2.upper_pixels_rows(chroms[0].name, 40) is None
    lib2.prefetch(chroms[0].name)
    assert lib2._cache[0] == chroms[0].name
    # loud failures
    b1 = np.repeat(np.arange(40), 3)
    b2 = b1 + np.tile(np.arange(3), 40)
    for cnt, what in ((np.full(120, 2.5), "non-integer"), (np.full(120, 2**31 + 5, np.int64), "outside int32")):
        p = str(tmp_path / ("odd_%s.cool" % what[:3]))
        _mini_cool(p, b1, b2, cnt)
        with pytest.raises(ValueError, match=what):
            coolio.open_map(p).upper_pixels_rows("chrZ", 8)
    b2l = b2.copy()
    b2l[30] = b1[30] - 2
    p = str(tmp_path / "lower.cool")
    _mini_cool(p, b1, b2l, np.ones(120, np.int32))
    with pytest.raises(ValueError, match="below the diagonal"):
        coolio.open_map(p).upper_pixels_rows("chrZ", 8)


def test_scoring_rows_leave_out_the_far_pixels_valid_does_not_need(tmp_path, monkeypatch):
    """The scoring path asks H5Cool for rows with the balancing weights: pixels beyond the band stay only where they
    make a bin valid that no nearer pixel does; the band sections are the full blob's, `depth` still gets everything."""
    from peakachu_b200 import rowpack, shard
    from tests.test_rowpack import _valid_mask
    chroms = _genome()
    path = str(tmp_path / "t.cool")
    h5write.write_cool(path, chroms, 10000, trans=_trans(chroms, np.random.default_rng(2)), chunk=700)
    lib = coolio.open_map(path)
    seen = 0
    for c in chroms:
        nd = 25
        w = lib.weights(c.name, "weight")
        full = lib.upper_pixels_rows(c.name, nd)
        for weights in (w, None):
            slim = lib.upper_pixels_rows(c.name, nd, scoring_weights=weights)
            hf, hs = rowpack.header(full), rowpack.header(slim)
            assert hs["n_far"] <= hf["n_far"] and hs["nnz_band"] == hf["nnz_band"]
            assert np.array_equal(full[hf["off_bits"]:hf["off_far_off"]], slim[hs["off_bits"]:hs["off_far_off"]])
            assert np.array_equal(_valid_mask(slim, weights), _valid_mask(full, weights))
            seen += hf["n_far"] - hs["n_far"]
        # what score_units / Chromosome.from_map hand to the engine
        monkeypatch.setenv("PEAKACHU_B200_SLIM_FAR", "1")
        enc, blob, _, _, size = shard._unit_columns(lib, c.name, nd, None, scoring_weights=w)
        assert np.array_equal(blob, lib.upper_pixels_rows(c.name, nd, scoring_weights=w)) and size == blob.size
        monkeypatch.setenv("PEAKACHU_B200_SLIM_FAR", "0")
        assert np.array_equal(shard._unit_columns(lib, c.name, nd, None, scoring_weights=w)[1], full)
        assert np.array_equal(shard._unit_columns(lib, c.name, nd, None)[1], full)          # `depth`, tools: everything
    assert seen > 0
