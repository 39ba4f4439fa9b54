"""``peakachu depth`` (calculate_depth.py): total intra-chromosomal contacts and the matching
pre-trained model. The per-chromosome sum ``np.triu(raw, k=min_dis // binsize).sum()``
(calculate_depth.py:25-28) is a device reduction over the uploaded pixel columns
(``pk_chrom_depth``), SURVEY.md section 8(f) row 4. ``.hic`` input is out of scope."""
import ctypes as C

import numpy as np

from . import _lib, coolio


def chromosome_depth(pixels, n_bins, min_dis_bins=0, device=0):
    """Sum of the counts of one chromosome's upper-triangle pixels with bin2 - bin1 >= min_dis_bins.
    ``pixels``: a packed-rows blob, (bin1_offset, bin2 - bin1 uint16, count uint16), (bin1_offset, bin2,
    count) or (bin1, bin2, count) as the ``coolio`` readers return them."""
    L = _lib.lib()
    _lib.require_device()
    h = C.c_void_p()
    _lib.check(L.pk_chrom_create(device, int(n_bins), 5, 6, 1, 0, None, C.byref(h)))
    try:
        if isinstance(pixels, np.ndarray):                     # packed pixel rows (rowpack)
            blob = _lib.as_c(pixels, np.uint8)
            _lib.check(L.pk_chrom_upload_rows(h, _lib.ptr(blob), blob.size, None, _lib.PK_MEM_HOST))
            tot = C.c_int64()
            _lib.check(L.pk_chrom_depth(h, int(min_dis_bins), C.byref(tot)))
            return int(tot.value)
        a, b, c = pixels
        if np.asarray(b).dtype == np.uint16:
            a, b, c = _lib.as_c(a, np.int64), _lib.as_c(b, np.uint16), _lib.as_c(c, np.uint16)
            _lib.check(L.pk_chrom_upload_csr16(h, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), b.size, None, _lib.PK_MEM_HOST))
        elif np.asarray(a).size == int(n_bins) + 1 and np.asarray(b).size != int(n_bins) + 1:
            a, b, c = _lib.as_c(a, np.int64), _lib.as_c(b, np.int32), _lib.as_c(c, np.int32)
            _lib.check(L.pk_chrom_upload_csr(h, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), b.size, None, _lib.PK_MEM_HOST))
        else:
            a, b, c = (_lib.as_c(v, np.int32) for v in (a, b, c))
            _lib.check(L.pk_chrom_upload_pixels(h, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), a.size, None, _lib.PK_MEM_HOST))
        tot = C.c_int64()
        _lib.check(L.pk_chrom_depth(h, int(min_dis_bins), C.byref(tot)))
        return int(tot.value)
    finally:
        L.pk_chrom_destroy(h)


def match_pretrained_models(v, platform="Hi-C"):
    """calculate_depth.py:46-68."""
    arr = [5000000, 10000000, 30000000, 50000000, 100000000, 150000000, 200000000, 250000000, 300000000,
           350000000, 400000000, 450000000, 500000000, 550000000, 600000000, 650000000, 700000000, 750000000,
           800000000, 850000000, 900000000, 1000000000, 1200000000, 1400000000, 1600000000, 1800000000,
           2000000000]
    idx = int(np.argmin(np.abs(v - np.r_[arr])))
    if arr[idx] >= 1000000000:
        return "{0:.2g} billion".format(arr[idx] / 1000000000)
    return "{0} million".format(arr[idx] // 1000000)


def main(args):
    Lib = coolio.open_map(args.path)
    device = getattr(args, "device", None) or 0
    if getattr(Lib, "chrom_lengths", None) is None:
        raise ValueError("%s has no chroms/length column: `depth` scales by the genome size "
                         "(calculate_depth.py:46)" % args.path)
    genome_size = int(np.sum(Lib.chrom_lengths))
    mindis = args.min_dis // Lib.binsize                                   # calculate_depth.py:22
    totals = 0
    for k in Lib.chromnames:
        print(k)
        pixels = Lib.upper_pixels_rows(k, 0) if hasattr(Lib, "upper_pixels_rows") else None
        if pixels is None:
            pixels = Lib.upper_pixels_csr16(k) if hasattr(Lib, "upper_pixels_csr16") else None
        if pixels is None:
            pixels = Lib.upper_pixels(k)
        totals += chromosome_depth(pixels, Lib.nbins(k), mindis, device)
    print("num of intra reads in your data:", totals)
    matched_read_num = 3031042417 / genome_size * totals                   # calculate_depth.py:42
    print("num of intra reads in a human with matched sequencing coverage:", int(matched_read_num))
    print("suggested model:", match_pretrained_models(matched_read_num))
    return totals
