"""Host-side mirror of the reference's ``peakachu.scoreUtils`` (scoreUtils.py:9-135).

``Chromosome`` keeps the reference's constructor signature, public attributes
(``exp_arr``, ``background``, ``ridx``, ``cidx``, ``M``) and methods
(``score(thre) -> (prob_csr, value_csr)``, ``writeBed(outfil, prob_csr, raw_csr)``),
but every numeric step runs in the CUDA library through the C ABI
(``include/peakachu_b200.h``). numpy arrays / scipy matrices are only containers
at the boundary; there is no CPU implementation of the path in this package.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .forest import FlatForest, flatten_forest


class DeviceForest:
    """A forest packed into device node tables (pk_forest_create)."""

    _cache: dict = {}

    def __init__(self, flat: FlatForest, device: int = 0):
        L = _lib.lib()
        self.flat, self.device = flat, device
        self.handle = C.c_void_p()
        no = _lib.as_c(flat.node_offset, np.int64)
        fe = _lib.as_c(flat.feature, np.int32)
        th = _lib.as_c(flat.threshold, np.float64)
        le = _lib.as_c(flat.left, np.int32)
        ri = _lib.as_c(flat.right, np.int32)
        ml = _lib.as_c(flat.missing_left, np.uint8)
        p1 = _lib.as_c(flat.leaf_p1, np.float64)
        _lib.check(L.pk_forest_create(device, flat.n_trees, flat.n_features, _lib.ptr(no, _lib.c_i64p),
                                      _lib.ptr(fe, _lib.c_i32p), _lib.ptr(th, _lib.c_f64p),
                                      _lib.ptr(le, _lib.c_i32p), _lib.ptr(ri, _lib.c_i32p),
                                      _lib.ptr(ml, _lib.c_u8p), _lib.ptr(p1, _lib.c_f64p),
                                      C.byref(self.handle)))

    @property
    def width(self):
        return self.flat.width

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            _lib.lib().pk_forest_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def of(cls, model, device: int = 0) -> "DeviceForest":
        """DeviceForest for a sklearn model / FlatForest / DeviceForest, cached per
        (object, device) so score_genome uploads the forest once."""
        if isinstance(model, DeviceForest):
            if model.device != device:
                return cls.of(model.flat, device)
            return model
        key = (id(model), device)
        hit = cls._cache.get(key)
        if hit is not None and hit[0] is model:
            return hit[1]
        flat = model if isinstance(model, FlatForest) else flatten_forest(model)
        df = cls(flat, device)
        cls._cache[key] = (model, df)
        return df


def _upper_pixels_from_csr(raw_M):
    """Upper-triangle (bin1, bin2, count) of a symmetric scipy matrix of counts."""
    from scipy import sparse
    up = sparse.triu(raw_M, k=0).tocoo()
    cnt = np.asarray(up.data)
    icnt = np.rint(cnt).astype(np.int32)
    if not np.array_equal(icnt, cnt):
        raise ValueError("raw matrix holds non-integer counts; the CUDA path stores int32 counts")
    return up.row.astype(np.int32), up.col.astype(np.int32), icnt


def _check_balanced_matrix(M, b1, b2, cnt, weights, sample=4096):
    """The device recomputes balanced values as ``(w[row] * w[col]) * count``; a caller whose ``M`` was balanced
    some other way (another weight column, a divisive one, a scaled matrix) would silently get different
    numbers. Compare a sample of ``M``'s stored pixels with that product and refuse a mismatch."""
    if b1.size == 0:
        return
    pick = np.unique(np.linspace(0, b1.size - 1, num=min(sample, b1.size)).astype(np.int64))
    r, c = b1[pick], b2[pick]
    have = np.asarray(M[r, c]).ravel().astype(np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        want = (weights[r] * weights[c]) * cnt[pick].astype(np.float64)
    same = (have == want) | (~np.isfinite(want) & (~np.isfinite(have) | (have == 0)))     # NaN pixels may be stored or dropped
    if not np.all(same):
        k = int(np.flatnonzero(~same)[0])
        raise ValueError("Chromosome: M[%d, %d] = %r but (weights[%d] * weights[%d]) * raw_M[%d, %d] = %r; the CUDA path "
                         "balances raw_M with `weights` itself, so M must be that product (cooler's balance=<the same "
                         "column>). Pass the weights that made M, or use weights=None with raw counts."
                         % (r[k], c[k], have[k], r[k], c[k], r[k], c[k], want[k]))


class Chromosome:
    """Drop-in for ``peakachu.scoreUtils.Chromosome`` (scoreUtils.py:9-38).

    Reference call sites: score_chromosome.py:45-48,51-54 and score_genome.py:58-61,64-67.
    ``M`` / ``raw_M`` are scipy CSR matrices as there; in balanced mode the balanced
    values are recomputed on the device from ``raw_M`` and ``weights`` as
    ``(w[row] * w[col]) * count`` (what ``cooler`` yields); ``M`` gives the shape and is
    checked against that product on a sample of pixels (a differently balanced ``M`` is
    refused rather than silently ignored). Use :meth:`from_pixels` to skip building matrices.
    """

    def __init__(self, M, model, raw_M=None, weights=None, lower=6, upper=300,
                 cname="chrm", res=10000, width=5, device=0, stream=None):
        if raw_M is None:
            raw_M = M
        if weights is None and M is not raw_M:
            raise NotImplementedError(
                "weights=None with M is not raw_M is the reference's .hic (KR/NONE via straw) branch, "
                "which is outside this path")
        b1, b2, cnt = _upper_pixels_from_csr(raw_M)
        if weights is not None and M is not raw_M:
            _check_balanced_matrix(M, b1, b2, cnt, np.asarray(weights, dtype=np.float64))
        self._init(b1, b2, cnt, weights, int(M.shape[0]), model, lower, upper, cname, res, width, device, stream,
                   sorted_pixels=None)

    @classmethod
    def from_pixels(cls, bin1, bin2, count, weights, n_bins, model, lower=6, upper=300,
                    cname="chrm", res=10000, width=5, device=0, stream=None, sorted_pixels=None,
                    first_tile=None, score_stream=None):
        """Build from cooler-style upper-triangle pixel columns (chromosome-local bin
        ids) and the weight column (None = raw mode). ``sorted_pixels=True`` promises
        cooler order (sorted by bin1, then bin2; the device verifies it), ``None``
        checks on the host, ``False`` takes the order-free scatter path."""
        self = cls.__new__(cls)
        self._init(bin1, bin2, count, weights, int(n_bins), model, lower, upper, cname, res, width, device, stream,
                   sorted_pixels=sorted_pixels, first_tile=first_tile, score_stream=score_stream)
        return self

    @classmethod
    def from_csr(cls, bin1_offset, bin2, count, weights, n_bins, model, lower=6, upper=300,
                 cname="chrm", res=10000, width=5, device=0, stream=None, first_tile=None, score_stream=None):
        """Build from cooler's CSR layout: ``indexes/bin1_offset`` of the chromosome
        (rebased to 0, int64[n_bins+1]) plus the ``bin2_id`` (chromosome-local) and
        ``count`` columns of its pixels. 8 bytes per pixel cross the bus instead of 12.
        Arrays may live in pinned host memory (e.g. views of pinned torch tensors)."""
        self = cls.__new__(cls)
        self._init(None, bin2, count, weights, int(n_bins), model, lower, upper, cname, res, width, device, stream,
                   bin1_offset=bin1_offset, first_tile=first_tile, score_stream=score_stream)
        return self

    @classmethod
    def from_csr16(cls, bin1_offset, bin2_delta, count, weights, n_bins, model, lower=6, upper=300,
                   cname="chrm", res=10000, width=5, device=0, stream=None, first_tile=None, score_stream=None):
        """``from_csr`` with narrow columns: ``bin2_delta = bin2 - bin1`` and ``count`` as uint16
        arrays (4 bytes per pixel cross the bus). Every pixel of the chromosome must be
        representable, see ``pk_chrom_upload_csr16`` in include/peakachu_b200.h;
        ``coolio.PKCool`` stores such columns when they are."""
        self = cls.__new__(cls)
        self._init(None, bin2_delta, count, weights, int(n_bins), model, lower, upper, cname, res, width, device,
                   stream, bin1_offset=bin1_offset, first_tile=first_tile, csr16=True, score_stream=score_stream)
        return self

    @classmethod
    def from_rows(cls, blob, weights, n_bins, model, lower=6, upper=300,
                  cname="chrm", res=10000, width=5, device=0, stream=None, first_tile=None, score_stream=None):
        """Build from packed pixel rows (``peakachu_b200.rowpack``; ``pk_chrom_upload_rows``): the
        chromosome crosses the bus in one copy of about 1.3 bytes per band pixel."""
        self = cls.__new__(cls)
        self._init(None, None, None, weights, int(n_bins), model, lower, upper, cname, res, width, device,
                   stream, first_tile=first_tile, score_stream=score_stream, rows=blob)
        return self

    @classmethod
    def from_map(cls, Lib, key, weights, model, lower=6, upper=300, cname="chrm", res=10000, width=5, device=0,
                 encoding=None, first_tile=None):
        """Build from an opened map (``coolio.open_map``), taking the chromosome's pixels in the most
        compact column format the reader offers: packed rows, uint16 columns, cooler's CSR columns."""
        from . import shard
        n = Lib.nbins(key)
        nd_need = min(upper, n - 2 * width) + 2 * width + 1
        enc, a, b, c, _ = shard._unit_columns(Lib, key, nd_need, encoding,
                                              scoring_weights=None if weights is None else np.asarray(weights, dtype=np.float64))
        kw = dict(lower=lower, upper=upper, cname=cname, res=res, width=width, device=device, first_tile=first_tile)
        if enc == _lib.PK_ENC_ROWS:
            return cls.from_rows(a, weights, n, model, **kw)
        if enc == _lib.PK_ENC_CSR16:
            return cls.from_csr16(a, b, c, weights, n, model, **kw)
        if enc == _lib.PK_ENC_CSR32:
            return cls.from_csr(a, b, c, weights, n, model, **kw)
        return cls.from_pixels(a, b, c, weights, n, model, sorted_pixels=True, **kw)

    # -- construction = upload + band + expected + candidates (scoreUtils.py:13-34) --
    def _init(self, b1, b2, cnt, weights, n, model, lower, upper, cname, res, width, device, stream,
              sorted_pixels=None, bin1_offset=None, first_tile=None, csr16=False, score_stream=None, rows=None):
        L = _lib.lib()
        _lib.require_device()
        self.chromname, self.r, self.w = cname, res, width
        self.model = model
        self.n = n
        self.device = device
        self.weights = None if weights is None else _lib.as_c(weights, np.float64)
        self._forest = None
        self._h = C.c_void_p()
        _lib.check(L.pk_chrom_create(device, n, width, lower, upper, 0 if weights is None else 1,
                                     C.c_void_p(stream or 0), C.byref(self._h)))
        if score_stream:
            # the scoring pass goes to its own stream (shard.score_units shares one among the
            # chromosomes in flight and gives `stream` a high priority)
            _lib.check(L.pk_chrom_set_score_stream(self._h, C.c_void_p(score_stream)))
        lo, up, el = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(L.pk_chrom_bounds(self._h, C.byref(lo), C.byref(up), C.byref(el)))
        self.lower, self.upper, self._exp_len = lo.value, up.value, el.value
        if rows is not None:
            blob = _lib.as_c(rows, np.uint8)
            self._keepalive = [blob]
            _lib.check(L.pk_chrom_upload_rows(self._h, _lib.ptr(blob), blob.size, _lib.ptr(self.weights),
                                              _lib.PK_MEM_HOST))
            self._after_upload(L, first_tile, n)
            return
        col_t = np.uint16 if csr16 else np.int32
        if csr16 and (np.asarray(b2).dtype != np.uint16 or np.asarray(cnt).dtype != np.uint16):
            raise TypeError("from_csr16 takes uint16 arrays (bin2 - bin1, count)")
        b2, cnt = _lib.as_c(b2, col_t), _lib.as_c(cnt, col_t)
        self._keepalive = [b2, cnt]          # uploads are asynchronous when the source is pinned
        if bin1_offset is not None:
            rp = _lib.as_c(bin1_offset, np.int64)
            self._keepalive.append(rp)
            if rp.size != n + 1:
                raise ValueError("bin1_offset must have n_bins + 1 entries")
            upload = L.pk_chrom_upload_csr16 if csr16 else L.pk_chrom_upload_csr
            _lib.check(upload(self._h, _lib.ptr(rp), _lib.ptr(b2), _lib.ptr(cnt), b2.size,
                              _lib.ptr(self.weights), _lib.PK_MEM_HOST))
        else:
            b1 = _lib.as_c(b1, np.int32)
            self._keepalive.append(b1)
            if sorted_pixels is None:
                sorted_pixels = bool(b1.size == 0 or (np.all(b1[1:] >= b1[:-1]) and np.all(b1 <= b2)))
            mem = _lib.PK_MEM_HOST | (_lib.PK_PIXELS_SORTED if sorted_pixels else 0)
            _lib.check(L.pk_chrom_upload_pixels(self._h, _lib.ptr(b1), _lib.ptr(b2), _lib.ptr(cnt), b1.size,
                                                _lib.ptr(self.weights), mem))
        self._after_upload(L, first_tile, n)

    def set_poisson_weights(self, weights):
        """Weights of the Poisson filter alone (``pk_chrom_set_poisson_weights``): the raw values of a
        ``divisive_weights`` column, whose reciprocals balanced the pixels. Re-runs the candidate scan."""
        wp = _lib.as_c(weights, np.float64)
        if wp.size != self.n:
            raise ValueError("poisson weights must have n_bins entries")
        self._keepalive.append(wp)
        L = _lib.lib()
        _lib.check(L.pk_chrom_set_poisson_weights(self._h, _lib.ptr(wp), _lib.PK_MEM_HOST))
        _lib.check(L.pk_chrom_find_candidates(self._h, 0, self.n, None))
        self._ncand = self._cand = None

    def _fit_expected_host(self, L):
        """PEAKACHU_B200_EXPECTED=host: the reference's own fit (utils.py:159-176) with the installed
        scikit-learn on the per-distance means the device computed, handed back with pk_chrom_set_expected."""
        from sklearn.isotonic import IsotonicRegression
        s = np.zeros(self._exp_len, dtype=np.float64)
        cnt = np.zeros(self._exp_len, dtype=np.int64)
        _lib.check(L.pk_chrom_diag_sums(self._h, _lib.ptr(s, _lib.c_f64p), _lib.ptr(cnt, _lib.c_i64p)))
        exp = np.zeros(self._exp_len)
        big = cnt > 10
        exp[big] = s[big] / cnt[big]
        d = np.where(exp > 0)[0]
        IR = IsotonicRegression(increasing=False, out_of_bounds="clip")
        IR.fit(d, exp[d])
        exp = np.ascontiguousarray(IR.predict(list(range(self._exp_len))), dtype=np.float64)
        _lib.check(L.pk_chrom_set_expected(self._h, _lib.ptr(exp, _lib.c_f64p), _lib.ptr(exp, _lib.c_f64p)))

    def _after_upload(self, L, first_tile, n):
        if _lib.expected_mode() == "host":
            self._fit_expected_host(L)
        else:
            _lib.check(L.pk_chrom_fit_expected(self._h))
        self._exp = None
        # asynchronous: the candidate count is read back only when somebody asks for it
        ra, rb = first_tile if first_tile is not None else (0, n)     # band row tile (multi-GPU seam)
        _lib.check(L.pk_chrom_find_candidates(self._h, int(ra), int(rb), None))
        self._ncand = None
        self._cand = None
        self.M = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().pk_chrom_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference attributes ------------------------------------------------------
    @property
    def n_candidates(self):
        if self._ncand is None:
            got = C.c_int64()
            _lib.check(_lib.lib().pk_chrom_candidates(self._h, None, None, 0, C.byref(got)))
            self._ncand = got.value
        return self._ncand

    @property
    def exp_arr(self):
        if self._exp is None:
            e = np.empty(self._exp_len, dtype=np.float64)
            _lib.check(_lib.lib().pk_chrom_get_expected(self._h, _lib.ptr(e, _lib.c_f64p)))
            self._exp = e
        return self._exp

    @property
    def background(self):
        return self.exp_arr

    def _candidates(self):
        if self._cand is None:
            n = self.n_candidates
            x, y = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
            got = C.c_int64()
            _lib.check(_lib.lib().pk_chrom_candidates(self._h, _lib.ptr(x, _lib.c_i32p), _lib.ptr(y, _lib.c_i32p),
                                                      n, C.byref(got)))
            self._cand = (x.astype(np.int64), y.astype(np.int64))
        return self._cand

    @property
    def ridx(self):
        return self._candidates()[0]

    @property
    def cidx(self):
        return self._candidates()[1]

    # -- parity tap: getwindow over the current candidates (scoreUtils.py:70-93) ------
    def window_features(self, want64=False):
        """(keep mask, float32 features[, float64 features]) for every candidate, in
        reference order; rows of rejected candidates are zero."""
        n, F = self.n_candidates, (2 * self.w + 1) ** 2
        keep = np.zeros(n, dtype=np.uint8)
        f32 = np.zeros((n, F), dtype=np.float32)
        f64 = np.zeros((n, F), dtype=np.float64) if want64 else None
        _lib.check(_lib.lib().pk_chrom_features(self._h, _lib.ptr(keep, _lib.c_u8p), _lib.ptr(f32, _lib.c_f32p),
                                                _lib.ptr(f64, _lib.c_f64p) if want64 else None, n))
        return (keep.astype(bool), f32, f64) if want64 else (keep.astype(bool), f32)

    def fused_window_features(self):
        """(keep mask, float32 features) as the product kernel (k_score_fused) builds them in
        shared memory: the parity tap inside the kernel ``score`` runs (pk_chrom_fused_features)."""
        if self._forest is None:
            self._forest = DeviceForest.of(self.model, self.device)
        n, F = self.n_candidates, (2 * self.w + 1) ** 2
        keep = np.zeros(n, dtype=np.uint8)
        f32 = np.zeros((n, F), dtype=np.float32)
        _lib.check(_lib.lib().pk_chrom_fused_features(self._h, self._forest.handle, _lib.ptr(keep, _lib.c_u8p),
                                                      _lib.ptr(f32, _lib.c_f32p), n))
        return keep.astype(bool), f32

    # -- scoring (scoreUtils.py:95-125) -----------------------------------------------
    def score_records(self, thre=0.5, with_batches=False):
        """(x, y, prob, value) numpy arrays sorted by (x, y); ``with_batches`` adds the records' reference
        batch ids and the surviving windows per batch (the row-tile seam: scoreUtils.py:104-108 is applied
        by the caller on the sums over the tiles)."""
        L = _lib.lib()
        if self._forest is None:
            self._forest = DeviceForest.of(self.model, self.device)
        _lib.check(L.pk_chrom_score(self._h, self._forest.handle, float(thre)))
        nrec, ncand, nwin = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(L.pk_chrom_result_count(self._h, C.byref(nrec), C.byref(ncand), C.byref(nwin)))
        n = nrec.value
        self.n_windows = nwin.value
        self._ncand = ncand.value
        x, y = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        p, v = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.float64)
        b = np.empty(n, dtype=np.int32) if with_batches else None
        _lib.check(L.pk_chrom_fetch_results(self._h, _lib.ptr(x), _lib.ptr(y), _lib.ptr(p), _lib.ptr(v), _lib.ptr(b),
                                            n, _lib.PK_MEM_HOST))
        if not with_batches:
            return x, y, p, v
        nb = C.c_int64()
        bw = np.zeros(max(1, self._ncand // 100000 + 2), dtype=np.int64)
        _lib.check(L.pk_chrom_batch_windows(self._h, _lib.ptr(bw, _lib.c_i64p), bw.size, C.byref(nb)))
        return x, y, p, v, b, bw[:nb.value]

    def score(self, thre=0.5):
        from scipy import sparse
        print("scoring matrix {}".format(self.chromname))
        print("number of candidates {}".format(self.n_candidates))
        x, y, p, v = self.score_records(thre)
        shape = (self.n, self.n)
        result = sparse.csr_matrix((p, (x, y)), shape=shape)
        self.M = sparse.csr_matrix((v, (x, y)), shape=shape) if x.size else result
        return result, self.M

    def stage_ms(self):
        ms = np.zeros(8, dtype=np.float32)
        _lib.check(_lib.lib().pk_chrom_stage_ms(self._h, _lib.ptr(ms, _lib.c_f32p)))
        return dict(zip(("band_build", "diag_sums", "expected_fit", "candidate_scan", "features", "forest",
                         "emit"), ms.tolist()))

    # -- output (scoreUtils.py:127-135) ------------------------------------------------
    def writeBed(self, outfil, prob_csr, raw_csr):
        r, c = prob_csr.nonzero()
        if r.size:
            pv = np.asarray(prob_csr[r, c]).ravel()
            rv = np.asarray(raw_csr[r, c]).ravel()
        else:
            pv = rv = np.zeros(0)
        with open(outfil, "a") as out:
            out.write(format_bedpe(self.chromname, self.r, r, c, pv, rv))


def format_bedpe(chromname, res, r, c, prob, value) -> str:
    """Rows exactly as scoreUtils.py:131-135 prints them: ints from int32 bin index
    times resolution, floats through str(numpy.float64) (shortest round-trip repr).
    Formatted by the library (pk_format_bedpe); host-only code, no device needed."""
    n = len(r)
    if n == 0:
        return ""
    L = _lib.lib()
    x, y = _lib.as_c(r, np.int32), _lib.as_c(c, np.int32)
    p, v = _lib.as_c(prob, np.float64), _lib.as_c(value, np.float64)
    name = chromname.encode()
    cap = n * (2 * len(name) + 4 * 21 + 2 * 26 + 8)
    buf = C.create_string_buffer(cap)
    written = C.c_int64()
    _lib.check(L.pk_format_bedpe(name, int(res), _lib.ptr(x, _lib.c_i32p), _lib.ptr(y, _lib.c_i32p),
                                 _lib.ptr(p, _lib.c_f64p), _lib.ptr(v, _lib.c_f64p), n, buf, cap, C.byref(written)))
    return buf.raw[:written.value].decode()
