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
