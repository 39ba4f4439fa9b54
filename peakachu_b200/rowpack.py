"""Packed pixel rows: the wire format of ``pk_chrom_upload_rows`` (include/peakachu_b200.h).

The scoring path reads a chromosome's pixels exactly once, to build the count band
``band[d][x]`` for d < upper + 2w + 1 (``scoreUtils.py:29-33``). cooler's columns spend
8 bytes (bin2 int32 + count int32; 16 in the file's own int64 / int32 layout) on a pixel
whose information is "row x has a pixel at distance d, with a small count". A packed row
holds the same pixels as

* a presence bitmap over the first ``nd_enc`` distances (``words_per_row`` uint32 per row),
* one byte per present pixel, in distance order (255 = "escape": the count is in a side list),
* the side list of escaped pixels (x, d, count) for counts >= 255 (mostly the main diagonal),
* the pixels farther out than ``nd_enc`` as plain CSR columns: they never enter the band, but
  the ``valid`` mask of ``utils.calculate_expected`` (``utils.py:146-156``) and ``peakachu
  depth`` (``calculate_depth.py:25-28``) look at every pixel of the chromosome.

All of it sits in ONE contiguous blob with a self-describing header, so that a chromosome crosses
the bus in a single copy: about 1.3 bytes per band pixel on a dense map instead of 8.
Encoding is lossless; ``unpack_rows`` restores the CSR columns (used by the tests).

Layout (little endian, every section 16-byte aligned; offsets are from the start of the blob):

    header   int64[16]: magic, n_bins, nd_enc, words_per_row, nnz_band, n_esc, n_far,
                        off_bits, off_cnt_off, off_cnt8, off_esc, off_far_off, off_far_b2,
                        off_far_cnt, total_bytes, 0
    bits     uint32[n_bins][words_per_row]
    cnt_off  uint32[n_bins + 1]         first byte of each row in cnt8
    cnt8     uint8[nnz_band]
    esc      int32[3][n_esc]            x | d | count, sorted by (x, d)
    far_off  int64[n_bins + 1]
    far_b2   int32[n_far]               chromosome-local bin2
    far_cnt  int32[n_far]
"""
from __future__ import annotations

import numpy as np

MAGIC = 0x31524B50          # "PKR1"
HEADER_WORDS = 16


def _align(v, a=16):
    return (v + a - 1) // a * a


def pack_rows(bin1_offset, bin2, count, n_bins: int, nd_enc: int) -> np.ndarray:
    """CSR pixel columns of one chromosome (cooler order: rows ascending, bin2 ascending inside a
    row, upper triangle, chromosome-local ids) -> packed blob (uint8 array). Duplicate pixels are
    summed, as ``utils.tocsr`` would (``utils.py:10-15``); zero counts are dropped."""
    rp = np.ascontiguousarray(bin1_offset, dtype=np.int64)
    b2 = np.ascontiguousarray(bin2, dtype=np.int64)
    cnt = np.ascontiguousarray(count, dtype=np.int64)
    n = int(n_bins)
    if rp.size != n + 1:
        raise ValueError("bin1_offset must have n_bins + 1 entries")
    nd_enc = int(nd_enc)
    if nd_enc < 1:
        raise ValueError("nd_enc must be positive")
    W = (nd_enc + 31) // 32
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    d = b2 - rows
    if d.size and (d.min() < 0 or b2.max() >= n):
        raise ValueError("pixels outside the upper triangle of the chromosome")
    if cnt.size and cnt.min() < 0:
        raise ValueError("negative pixel counts")
    if d.size > 1:
        key = rows * n + b2
        if np.any(np.diff(key) < 0):
            raise ValueError("pixels are not in cooler order (bin1, then bin2)")
        dup = np.diff(key) == 0
        if dup.any():                                    # sum duplicates (utils.tocsr semantics)
            first = np.concatenate([[True], ~dup])
            cnt = np.add.reduceat(cnt, np.nonzero(first)[0])
            rows, b2, d = rows[first], b2[first], d[first]
    keep = cnt > 0
    if not keep.all():
        rows, b2, d, cnt = rows[keep], b2[keep], d[keep], cnt[keep]
    if cnt.size and cnt.max() > np.iinfo(np.int32).max:
        raise ValueError("pixel counts beyond int32")
    inb = d < nd_enc
    r_in, d_in, c_in = rows[inb], d[inb], cnt[inb]
    if r_in.size >= 2 ** 32:
        raise ValueError("more than 2^32 band pixels in one chromosome")
    # presence bits: bits are unique after de-duplication, so a sum is an OR (exact in float64: < 2^32)
    bits = np.bincount(r_in * W + (d_in >> 5), weights=np.left_shift(1, d_in & 31).astype(np.float64),
                       minlength=n * W).astype(np.uint32)
    cnt_off = np.concatenate([[0], np.cumsum(np.bincount(r_in, minlength=n))]).astype(np.uint32)
    cnt8 = np.minimum(c_in, 255).astype(np.uint8)
    e = c_in >= 255
    esc = np.stack([r_in[e], d_in[e], c_in[e]]).astype(np.int32) if e.any() else np.zeros((3, 0), np.int32)
    far = ~inb
    far_off = np.concatenate([[0], np.cumsum(np.bincount(rows[far], minlength=n))]).astype(np.int64)
    far_b2, far_cnt = b2[far].astype(np.int32), cnt[far].astype(np.int32)

    sections = [bits, cnt_off, cnt8, esc, far_off, far_b2, far_cnt]
    offs, at = [], HEADER_WORDS * 8
    for a in sections:
        at = _align(at)
        offs.append(at)
        at += a.nbytes
    total = _align(at)
    blob = np.zeros(total, dtype=np.uint8)
    head = np.array([MAGIC, n, nd_enc, W, r_in.size, esc.shape[1], far_b2.size] + offs + [total, 0], dtype=np.int64)
    assert head.size == HEADER_WORDS
    blob[:head.nbytes] = head.view(np.uint8)
    for a, o in zip(sections, offs):
        if a.nbytes:
            blob[o:o + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    return blob


def header(blob) -> dict:
    h = np.frombuffer(memoryview(blob)[:HEADER_WORDS * 8], dtype=np.int64)
    if int(h[0]) != MAGIC:
        raise ValueError("not a packed-rows blob")
    names = ("magic", "n_bins", "nd_enc", "words_per_row", "nnz_band", "n_esc", "n_far", "off_bits", "off_cnt_off",
             "off_cnt8", "off_esc", "off_far_off", "off_far_b2", "off_far_cnt", "total_bytes")
    return {k: int(v) for k, v in zip(names, h)}


def unpack_rows(blob):
    """Inverse of ``pack_rows``: (bin1_offset int64[n+1], bin2 int32, count int32) in cooler order."""
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    h = header(blob)
    n, W = h["n_bins"], h["words_per_row"]

    def sec(name, dtype, count):
        return np.frombuffer(blob, dtype=dtype, count=count, offset=h[name])
    bits = sec("off_bits", np.uint32, n * W).reshape(n, W)
    cnt8 = sec("off_cnt8", np.uint8, h["nnz_band"])
    esc = sec("off_esc", np.int32, 3 * h["n_esc"]).reshape(3, h["n_esc"])
    far_off = sec("off_far_off", np.int64, n + 1)
    far_b2 = sec("off_far_b2", np.int32, h["n_far"])
    far_cnt = sec("off_far_cnt", np.int32, h["n_far"])
    present = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")        # [n][32 W]
    r_in, d_in = np.nonzero(present)
    c_in = cnt8.astype(np.int32)
    if h["n_esc"]:
        pos = np.searchsorted(r_in.astype(np.int64) * (32 * W) + d_in, esc[0].astype(np.int64) * (32 * W) + esc[1])
        c_in[pos] = esc[2]
    far_rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(far_off))
    rows = np.concatenate([r_in.astype(np.int64), far_rows])
    b2 = np.concatenate([(r_in + d_in).astype(np.int64), far_b2.astype(np.int64)])
    cc = np.concatenate([c_in, far_cnt])
    order = np.lexsort((b2, rows))
    rows, b2, cc = rows[order], b2[order], cc[order]
    rp = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int64)
    return rp, b2.astype(np.int32), cc.astype(np.int32)


def pack_rows_native(bin1_offset, bin2, count, n_bins: int, nd_enc: int, bin2_base: int = 0, n_threads: int = 0,
                     valid_only_far: bool = False, weights=None) -> np.ndarray:
    """``pack_rows`` by the library (``pk_rows_pack``: two threaded passes over the columns, no temporaries): the
    same blob, from the rows of one chromosome as a cooler file stores them -- ``bin2`` may hold genome-wide ids
    (``bin2_base`` = the chromosome's first bin is subtracted; pixels behind the chromosome are dropped) and keeps
    the file's integer width, ``count`` its type. ``valid_only_far``: of the pixels beyond ``nd_enc`` keep only those
    the scoring path needs -- the finite ones (``weights``: the balancing weights, None for raw counts) that make a
    bin valid which no pixel inside ``nd_enc`` makes valid (``utils.py:146-156``); ``depth`` needs them all."""
    import ctypes as C

    from . import _lib
    L = _lib.lib()
    rp = np.ascontiguousarray(bin1_offset, dtype=np.int64)
    if rp.size != int(n_bins) + 1:
        raise ValueError("bin1_offset must have n_bins + 1 entries")
    b2 = np.ascontiguousarray(bin2)
    if b2.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
        b2 = b2.astype(np.int64)
    cnt = np.ascontiguousarray(count)
    kind = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float64): 2}.get(cnt.dtype)
    if kind is None:
        if cnt.dtype.kind == "u" and cnt.dtype.itemsize == 8 and cnt.size and int(cnt.max()) > np.iinfo(np.int64).max:
            raise ValueError("pixel counts outside int32")
        cnt = cnt.astype(np.float64 if cnt.dtype.kind == "f" else np.int64)
        kind = 2 if cnt.dtype.kind == "f" else 1
    if b2.size != cnt.size or (rp.size and int(rp[-1]) > b2.size):
        raise ValueError("bin1_offset points behind the pixel columns")
    need = C.c_int64()
    w = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.size != int(n_bins):
            raise ValueError("weights must have n_bins entries")
    args = (_lib.ptr(rp, _lib.c_i64p), C.c_void_p(b2.ctypes.data), b2.dtype.itemsize, int(bin2_base),
            C.c_void_p(cnt.ctypes.data), kind, int(n_bins), int(nd_enc), 1 if valid_only_far else 0,
            C.c_void_p(w.ctypes.data) if w is not None else None)
    rc = L.pk_rows_pack(*args, None, 0, C.byref(need), int(n_threads))
    if rc != 0:
        raise ValueError(L.pk_last_error().decode("utf-8", "replace"))
    blob = np.empty(need.value, dtype=np.uint8)
    rc = L.pk_rows_pack(*args, C.c_void_p(blob.ctypes.data), blob.size, C.byref(need), int(n_threads))
    if rc != 0:
        raise ValueError(L.pk_last_error().decode("utf-8", "replace"))
    return blob
