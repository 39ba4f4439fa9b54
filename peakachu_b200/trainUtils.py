"""Training-set window features on the GPU: drop-in for ``peakachu.trainUtils.buildmatrix``
(trainUtils.py:12-44), SURVEY.md section 8(f) row 3.

The reference slices (2w+1)^2 windows around the given pixels out of the balanced matrix,
normalises them by an expected curve fitted over ``max|i-j| + 2w`` distances, Gaussian-filters
and min-max scales them -- the same arithmetic as ``Chromosome.getwindow`` with three
differences (no band trim, its own expected length, the extra mask ``y - x > w``). Here the
windows come from the same CUDA kernel as the scoring path's feature tap
(``pk_chrom_features_at``); fitting the forest stays with scikit-learn on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .scoreUtils import _upper_pixels_from_csr


def buildmatrix_from_pixels(bin1, bin2, count, weights, n_bins, coords, w=5, device=0):
    """``buildmatrix`` from cooler-style upper-triangle pixel columns (chromosome-local bin ids)
    and the weight column (None = raw counts). Returns a list of float64 feature vectors, one
    per pixel that survives the reference's filters, in input order -- or None when fewer than
    ten pixels pass the coordinate mask (trainUtils.py:24-25)."""
    L = _lib.lib()
    _lib.require_device()
    coords = np.r_[coords]
    xi, yi = coords[:, 0].astype(np.int64), coords[:, 1].astype(np.int64)
    n = int(n_bins)
    mask = (xi - w >= 0) & (yi + w + 1 <= n) & (yi - xi > w)                 # trainUtils.py:22
    xi, yi = xi[mask], yi[mask]
    if xi.size < 10:
        return None
    maxd = int(np.abs(xi - yi).max())
    maxdis = maxd + 2 * w                                                    # trainUtils.py:28
    b1, b2, cnt = (_lib.as_c(a, np.int32) for a in (bin1, bin2, count))
    wts = None if weights is None else _lib.as_c(weights, np.float64)
    h = C.c_void_p()
    # one distance more than the reference's expected curve covers: the far corner of the
    # farthest window lies at distance maxdis and must be a stored diagonal
    _lib.check(L.pk_chrom_create(device, n, w, w + 1, maxd + 1, 0 if wts is None else 1, None, C.byref(h)))
    try:
        lo, up, el = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(L.pk_chrom_bounds(h, C.byref(lo), C.byref(up), C.byref(el)))
        if el.value != maxdis + 2:
            raise RuntimeError("unexpected expected-curve length %d (wanted %d)" % (el.value, maxdis + 2))
        sorted_pixels = bool(b1.size == 0 or (np.all(b1[1:] >= b1[:-1]) and np.all(b1 <= b2)))
        mem = _lib.PK_MEM_HOST | (_lib.PK_PIXELS_SORTED if sorted_pixels else 0)
        _lib.check(L.pk_chrom_upload_pixels(h, _lib.ptr(b1), _lib.ptr(b2), _lib.ptr(cnt), b1.size, _lib.ptr(wts), mem))
        # utils.calculate_expected(Matrix, maxdis): the fit sees distances 0..maxdis only
        s = np.zeros(el.value, np.float64)
        k = np.zeros(el.value, np.int64)
        _lib.check(L.pk_chrom_diag_sums(h, _lib.ptr(s, _lib.c_f64p), _lib.ptr(k, _lib.c_i64p)))
        e = np.zeros(el.value, np.float64)
        _lib.check(L.pk_fit_expected(_lib.ptr(s, _lib.c_f64p), _lib.ptr(k, _lib.c_i64p), maxdis + 1, _lib.ptr(e, _lib.c_f64p)))
        e[maxdis + 1] = e[maxdis]                                            # never read by a window
        _lib.check(L.pk_chrom_set_expected(h, _lib.ptr(e, _lib.c_f64p), _lib.ptr(e, _lib.c_f64p)))
        x32, y32 = _lib.as_c(xi, np.int32), _lib.as_c(yi, np.int32)
        F = (2 * w + 1) ** 2
        keep = np.zeros(xi.size, np.uint8)
        f64 = np.zeros((xi.size, F), np.float64)
        _lib.check(L.pk_chrom_features_at(h, _lib.ptr(x32), _lib.ptr(y32), xi.size, _lib.ptr(keep, _lib.c_u8p), None,
                                          _lib.ptr(f64, _lib.c_f64p)))
    finally:
        L.pk_chrom_destroy(h)
    return [f64[i] for i in np.nonzero(keep)[0]]


def buildmatrix(Matrix, coords, w=5, raw_M=None, weights=None, device=0):
    """Reference signature plus what the CUDA path needs to rebuild the balanced values:
    ``raw_M`` (scipy matrix of counts) and ``weights`` -- as ``scoreUtils.Chromosome`` takes
    them. With both None, ``Matrix`` itself must hold raw counts (``--clr-weight-name raw``)."""
    counts = Matrix if raw_M is None else raw_M
    if raw_M is not None and weights is None:
        raise NotImplementedError("balanced values without a weight column (.hic) are not supported")
    b1, b2, cnt = _upper_pixels_from_csr(counts)
    return buildmatrix_from_pixels(b1, b2, cnt, weights, Matrix.shape[0], coords, w=w, device=device)
