"""Packed pixel rows (peakachu_b200/rowpack.py, the wire format of pk_chrom_upload_rows): lossless on
synthetic maps, with duplicates, escaped counts, far pixels and empty rows. CPU only."""
import numpy as np
import pytest

from peakachu_b200 import rowpack, synth


def _csr(ch):
    rp = np.searchsorted(ch.bin1, np.arange(ch.n + 1)).astype(np.int64)
    return rp, ch.bin2, ch.count


@pytest.mark.parametrize("n,nd,depth", [(400, 71, 300.0), (2000, 320, 300.0), (900, 97, 8.0), (64, 32, 2000.0)])
def test_round_trip(n, nd, depth):
    ch = synth.make_chromosome("c", n, seed=n, depth=depth, band=min(330, n))
    rp, b2, cnt = _csr(ch)
    blob = rowpack.pack_rows(rp, b2, cnt, n, nd)
    h = rowpack.header(blob)
    assert h["n_bins"] == n and h["nd_enc"] == nd and h["total_bytes"] == blob.size
    assert h["nnz_band"] == int(((b2 - ch.bin1) < nd).sum()) and h["n_far"] == int(((b2 - ch.bin1) >= nd).sum())
    assert h["n_esc"] == int(((cnt >= 255) & ((b2 - ch.bin1) < nd)).sum())
    rp2, b22, cnt2 = rowpack.unpack_rows(blob)
    assert np.array_equal(rp, rp2) and np.array_equal(b2, b22) and np.array_equal(cnt, cnt2)


def test_duplicates_are_summed_and_zero_counts_dropped():
    rp = np.array([0, 4, 4, 5], dtype=np.int64)
    b2 = np.array([0, 1, 1, 2, 2], dtype=np.int32)
    cnt = np.array([300, 2, 3, 0, 7], dtype=np.int32)
    rp2, b22, cnt2 = rowpack.unpack_rows(rowpack.pack_rows(rp, b2, cnt, 3, 2))
    assert rp2.tolist() == [0, 2, 2, 3] and b22.tolist() == [0, 1, 2] and cnt2.tolist() == [300, 5, 7]


def test_bad_input_is_refused():
    rp = np.array([0, 1, 2], dtype=np.int64)
    with pytest.raises(ValueError):
        rowpack.pack_rows(rp, np.array([1, 0]), np.array([1, 1]), 2, 8)         # below the diagonal
    with pytest.raises(ValueError):
        rowpack.pack_rows(rp, np.array([0, 5]), np.array([1, 1]), 2, 8)         # outside the chromosome
    with pytest.raises(ValueError):
        rowpack.pack_rows(np.array([0, 2, 2], dtype=np.int64), np.array([1, 0]), np.array([1, 1]), 2, 8)   # unsorted row
    with pytest.raises(ValueError):
        rowpack.header(np.zeros(128, np.uint8))


@pytest.mark.parametrize("n,nd,depth", [(400, 71, 300.0), (2000, 320, 300.0), (900, 97, 8.0), (64, 32, 2000.0), (1, 5, 50.0)])
@pytest.mark.parametrize("threads", [1, 3])
def test_native_packer_writes_the_same_blob(n, nd, depth, threads):
    """pk_rows_pack (threaded C passes over cooler's own columns) against the numpy packer, byte for byte: plain
    chromosome-local columns, and the same chromosome as rows of a genome-wide file -- int64 genome-wide bin2 ids
    with inter-chromosomal pixels at the end of every row, counts as int64 / float64, duplicates and zero counts."""
    ch = synth.make_chromosome("c", n, seed=n, depth=depth, band=min(330, n))
    rp, b2, cnt = _csr(ch)
    want = rowpack.pack_rows(rp, b2, cnt, n, nd)
    got = rowpack.pack_rows_native(rp, b2.astype(np.int32), cnt.astype(np.int32), n, nd, n_threads=threads)
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    # as stored in a genome-wide file: chromosome at bins [base, base + n), a few trans pixels behind every third row,
    # some pixels split in two (duplicates) and some zero counts in between
    rng = np.random.default_rng(n)
    base = 1000
    rows = np.repeat(np.arange(n), np.diff(rp))
    out_b2, out_cnt, out_rows = [], [], []
    for x in range(n):
        s, e = rp[x], rp[x + 1]
        rb2, rc = b2[s:e].astype(np.int64) + base, cnt[s:e].astype(np.int64)
        if rb2.size and x % 2 == 0:                       # split the first pixel of the row into two stored pixels
            k = int(rc[0]) // 2
            rb2 = np.concatenate([[rb2[0]], rb2]); rc = np.concatenate([[k], [rc[0] - k], rc[1:]])
        if x % 5 == 0 and x + 1 < n:                      # a stored zero (dropped)
            zb = base + x + 1
            if zb not in rb2:
                pos = int(np.searchsorted(rb2, zb))
                rb2 = np.insert(rb2, pos, zb); rc = np.insert(rc, pos, 0)
        if x % 3 == 0:                                    # inter-chromosomal pixels
            t = np.sort(rng.integers(base + n, base + n + 500, 3))
            rb2 = np.concatenate([rb2, t]); rc = np.concatenate([rc, [4, 5, 6]])
        out_b2.append(rb2); out_cnt.append(rc); out_rows.append(rb2.size)
    g_rp = np.concatenate([[0], np.cumsum(out_rows)]).astype(np.int64)
    g_b2, g_cnt = np.concatenate(out_b2), np.concatenate(out_cnt)
    for cast in (np.int64, np.float64, np.uint16 if g_cnt.max() < 65536 else np.int64):
        got = rowpack.pack_rows_native(g_rp, g_b2, g_cnt.astype(cast), n, nd, bin2_base=base, n_threads=threads)
        assert np.array_equal(got, want), cast


def test_native_packer_refuses_what_the_numpy_one_refuses():
    rp = np.array([0, 1, 2], dtype=np.int64)
    one = np.array([1, 1], dtype=np.int32)
    with pytest.raises(ValueError, match="below the diagonal"):
        rowpack.pack_rows_native(rp, np.array([1, 0], dtype=np.int32), one, 2, 8)
    with pytest.raises(ValueError, match="cooler order"):
        rowpack.pack_rows_native(np.array([0, 2, 2], dtype=np.int64), np.array([1, 0], dtype=np.int32), one, 2, 8)
    with pytest.raises(ValueError, match="cooler order"):            # a cis pixel behind an inter-chromosomal one
        rowpack.pack_rows_native(np.array([0, 2, 2], dtype=np.int64), np.array([7, 1], dtype=np.int32), one, 2, 8)
    with pytest.raises(ValueError, match="negative"):
        rowpack.pack_rows_native(rp, np.array([0, 1], dtype=np.int32), np.array([1, -1], dtype=np.int32), 2, 8)
    with pytest.raises(ValueError, match="outside int32"):
        rowpack.pack_rows_native(rp, np.array([0, 1], dtype=np.int32), np.array([1, 2**31], dtype=np.int64), 2, 8)
    with pytest.raises(ValueError, match="outside int32"):            # duplicates whose sum overflows
        rowpack.pack_rows_native(np.array([0, 2, 2], dtype=np.int64), np.array([0, 0], dtype=np.int32),
                                 np.array([2**30, 2**30], dtype=np.int64), 2, 8)
    with pytest.raises(ValueError, match="non-integer"):
        rowpack.pack_rows_native(rp, np.array([0, 1], dtype=np.int32), np.array([1.0, 2.5]), 2, 8)
    with pytest.raises(ValueError, match="n_bins"):
        rowpack.pack_rows_native(rp, np.array([0, 1], dtype=np.int32), one, 3, 8)
    # a trans-only row and an empty chromosome are fine
    blob = rowpack.pack_rows_native(rp, np.array([9, 1], dtype=np.int32), one, 2, 8)
    assert rowpack.unpack_rows(blob)[0].tolist() == [0, 0, 1]
    blob = rowpack.pack_rows_native(np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32), 0, 8)
    assert rowpack.header(blob)["n_bins"] == 0
