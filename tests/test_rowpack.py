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


def _valid_mask(blob, weights):
    """The `valid` mask the band build derives from a blob (k_band_rows / k_band_escapes, utils.py:146-156): a pixel
    with count > 0 and -- with weights -- a finite balanced value (w_x w_y) count makes both its bins valid."""
    rp, b2, cnt = rowpack.unpack_rows(blob)
    n = rp.size - 1
    x = np.repeat(np.arange(n), np.diff(rp))
    fin = cnt > 0
    if weights is not None:
        with np.errstate(all="ignore"):
            fin &= np.isfinite((weights[x] * weights[b2]) * cnt.astype(np.float64))
    valid = np.zeros(n, bool)
    valid[x[fin]] = True
    valid[b2[fin]] = True
    return valid


@pytest.mark.parametrize("seed", range(6))
def test_far_pixels_reduced_to_the_witnesses_of_valid(seed):
    """valid_only_far (what the scoring path uploads from a cooler file): the sections the band is built from are
    untouched, the far lists shrink to a subset, and the `valid` mask is the one the full blob gives -- with NaN,
    infinite, zero, negative and overflowing weights, bins whose only finite pixels are far, escaped counts, and
    in raw mode."""
    rng = np.random.default_rng(seed)
    n, nd = 260, 12
    rows_b2, rows_c = [], []
    sparse_rows = set(rng.choice(n, 60, replace=False).tolist())          # rows without pixels inside nd
    for x in range(n):
        near = np.arange(x, min(n, x + nd))
        near = near[(rng.random(near.size) < (0.0 if x in sparse_rows else 0.6)) & ~np.isin(near, list(sparse_rows))]
        far = np.arange(min(n, x + nd), n)
        far = far[rng.random(far.size) < 0.04]
        b2 = np.concatenate([near, far])
        rows_b2.append(b2)
        rows_c.append(rng.choice([1, 2, 3, 300, 7], b2.size))
    rp = np.concatenate([[0], np.cumsum([b.size for b in rows_b2])]).astype(np.int64)
    b2 = np.concatenate(rows_b2).astype(np.int32)
    cnt = np.concatenate(rows_c).astype(np.int32)
    w = rng.uniform(0.5, 1.5, n)
    w[rng.choice(n, 25, replace=False)] = np.nan
    w[rng.choice(n, 6, replace=False)] = np.inf
    w[rng.choice(n, 6, replace=False)] = 1e200                            # products of two overflow
    w[rng.choice(n, 6, replace=False)] = 0.0
    w[rng.choice(n, 6, replace=False)] = -1.0
    for weights in (w, None):
        full = rowpack.pack_rows_native(rp, b2, cnt, n, nd)
        assert np.array_equal(full, rowpack.pack_rows(rp, b2, cnt, n, nd))
        slim = rowpack.pack_rows_native(rp, b2, cnt, n, nd, valid_only_far=True, weights=weights, n_threads=1 + seed % 3)
        hf, hs = rowpack.header(full), rowpack.header(slim)
        for k in ("n_bins", "nd_enc", "words_per_row", "nnz_band", "n_esc"):
            assert hf[k] == hs[k]
        assert hs["n_far"] < hf["n_far"] and hs["total_bytes"] < hf["total_bytes"]
        assert np.array_equal(full[hf["off_bits"]:hf["off_far_off"]], slim[hs["off_bits"]:hs["off_far_off"]])
        assert np.array_equal(_valid_mask(slim, weights), _valid_mask(full, weights))
        # the kept far pixels are pixels of the full list, each finite and a witness of a bin nothing nearer makes valid
        rf, bf, cf = rowpack.unpack_rows(full)
        rs, bs, cs = rowpack.unpack_rows(slim)
        key = lambda r, b: set(zip(np.repeat(np.arange(n), np.diff(r)).tolist(), b.tolist()))
        assert key(rs, bs) <= key(rf, bf)
        xs = np.repeat(np.arange(n), np.diff(rs))
        far = (bs - xs) >= nd
        near_only = rowpack.pack_rows_native(rs, bs, np.where(far, 0, cs).astype(np.int32), n, nd)      # far counts zeroed: dropped
        near_valid = _valid_mask(near_only, weights)
        assert far.sum() == hs["n_far"] > 0 and np.all(~near_valid[xs[far]] | ~near_valid[bs[far]])
        assert not np.array_equal(near_valid, _valid_mask(full, weights))      # some bins are valid through far pixels only
    with pytest.raises(ValueError, match="n_bins"):
        rowpack.pack_rows_native(rp, b2, cnt, n, nd, valid_only_far=True, weights=w[:-1])
