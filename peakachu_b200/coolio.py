"""Contact-map ingestion for the scoring path.

The reference reads a ``.cool`` through four ``cooler`` calls
(``score_chromosome.py:33-44``, ``score_genome.py:28-61``)::

    Lib = cooler.Cooler(uri)
    Lib.chromnames
    Lib.matrix(balance=<name>|False, sparse=True).fetch(chrom)   # full symmetric COO
    Lib.bins().fetch(chrom)[<name>].values                       # float64[n]

``cooler``/``h5py`` are not installed in this image, so this module provides

* ``H5Cool`` -- a reader of real cooler files over the built-in HDF5 subset reader
  (``h5mini``), with PKCool's surface,
* ``PKCool`` -- an on-disk container (``.pkcool`` = numpy ``.npz``) holding the
  same columns a cooler file holds (``pixels/{bin1_id,bin2_id,count}`` with
  genome-wide bin ids, ``bins/<weight>``, ``chroms/{name,length}``,
  ``indexes/chrom_offset``), and
* ``Cooler`` -- a stand-in with exactly the four calls above, so the reference's
  unmodified ``main(args)`` functions run on a ``.pkcool`` when this module is
  injected as ``sys.modules['cooler']`` (tests only).

Balanced value of a pixel is defined here as ``(w[row] * w[col]) * count`` on the
full symmetric matrix, diagonal once -- cooler's ``bias1[row] * bias2[col] * data``.
cooler itself is absent, so this definition is the boundary's contract
("parity unpinned at the cooler boundary").

``open_map`` is what the product path calls: it returns the upper-triangle pixel
triplets and the weight vector per chromosome (what the CUDA band build
consumes), from a ``.pkcool`` or from a real ``.cool`` / ``.mcool::group`` URI.
"""
from __future__ import annotations

import os

import numpy as np


class PKCool:
    """Container with cooler's column layout. Pixels are intra-chromosomal,
    upper-triangle, sorted by (bin1_id, bin2_id), genome-wide bin ids."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        z = np.load(path, allow_pickle=False)
        self.path = path
        self._z = z
        self._cache = {}
        self.binsize = int(z["binsize"])
        self.chromnames = [str(s) for s in z["chrom_names"]]
        self.chrom_lengths = z["chrom_lengths"].astype(np.int64)      # bp
        self.chrom_offset = z["chrom_offset"].astype(np.int64)        # bins, len nchrom+1
        # The pixel columns (bin1_id, bin2_id, count; optionally delta16 / count16 = bin2 - bin1 and count as
        # uint16, written when every pixel is representable; optionally one blob of packed pixel rows per
        # chromosome, peakachu_b200.rowpack: what crosses the bus) are mapped on first use: a scoring run that
        # finds its packed rows never touches the plain columns, which are ten times the bytes.
        self.rows_nd = int(z["rows_nd"]) if "rows_nd" in z.files else 0
        self._row_keys = {int(k[len("rows_"):]): k for k in z.files if k.startswith("rows_") and k != "rows_nd"}
        self.weight_columns = {k[len("bins_"):]: z[k] for k in z.files if k.startswith("bins_")}
        # weight columns cooler would invert (column attribute `divisive_weights`, e.g. hic2cool's KR / VC)
        self.divisive = {str(s) for s in z["divisive_columns"]} if "divisive_columns" in z.files else set()

    def _col(self, key: str):
        """Column ``key`` of the container (None when absent). Members stored without compression (what ``write``
        produces) are memory-mapped in place -- no copy, no CRC pass; anything else is read through numpy."""
        if key in self._cache:
            return self._cache[key]
        a = None
        if key in self._z.files:
            a = self._map_member(key + ".npy")
            if a is None:
                a = self._z[key]
        self._cache[key] = a
        return a

    def _map_member(self, member: str):
        import struct
        import zipfile
        from numpy.lib import format as npf
        try:
            info = self._z.zip.getinfo(member)
            if info.compress_type != zipfile.ZIP_STORED:
                return None
            with open(self.path, "rb") as fh:
                fh.seek(info.header_offset)
                hdr = fh.read(30)
                if hdr[:4] != b"PK\x03\x04":
                    return None
                n_name, n_extra = struct.unpack("<HH", hdr[26:30])
                fh.seek(info.header_offset + 30 + n_name + n_extra)
                major, minor = npf.read_magic(fh)
                shape, fortran, dtype = (npf.read_array_header_1_0 if major == 1 else npf.read_array_header_2_0)(fh)
                if dtype.hasobject or fortran or len(shape) != 1:
                    return None
                off = fh.tell()
            if shape[0] == 0:
                return np.zeros(0, dtype=dtype)
            return np.memmap(self.path, dtype=dtype, mode="r", offset=off, shape=shape)
        except Exception:
            return None

    bin1_id = property(lambda self: self._col("bin1_id"))
    bin2_id = property(lambda self: self._col("bin2_id"))
    count = property(lambda self: self._col("count"))
    delta16 = property(lambda self: self._col("delta16"))
    count16 = property(lambda self: self._col("count16"))

    def _pix_range(self, i: int):
        """pixel range of chromosome i (bin1 sorted, intra-chromosomal only)"""
        if "_pix" not in self._cache:
            self._cache["_pix"] = (np.searchsorted(self.bin1_id, self.chrom_offset[:-1], side="left"),
                                   np.searchsorted(self.bin1_id, self.chrom_offset[1:], side="left"))
        lo, hi = self._cache["_pix"]
        return int(lo[i]), int(hi[i])

    # -- writer -----------------------------------------------------------
    @staticmethod
    def write(path: str, chroms, binsize: int, weight_name: str = "weight", rows_nd: int = 352,
              divisive: bool = False) -> None:
        """chroms: iterable of synth.SynthChrom-like objects
        (name, n, bin1, bin2, count, weights). ``rows_nd`` > 0 also stores every chromosome as packed
        pixel rows covering that many distances (enough for ``--upper 300`` with windows up to 25 x 25;
        a run that needs more distances falls back to the plain columns)."""
        from . import rowpack
        chroms = list(chroms)
        names = np.array([c.name for c in chroms])
        nb = np.array([c.n for c in chroms], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        b1 = np.concatenate([c.bin1.astype(np.int64) + off[i] for i, c in enumerate(chroms)])
        b2 = np.concatenate([c.bin2.astype(np.int64) + off[i] for i, c in enumerate(chroms)])
        cnt = np.concatenate([c.count for c in chroms]).astype(np.int32)
        w = np.concatenate([c.weights for c in chroms]).astype(np.float64)
        extra = {}
        if b1.size and int((b2 - b1).max()) <= 65535 and int(cnt.max()) <= 65535 and int(cnt.min()) >= 0:
            extra = dict(delta16=(b2 - b1).astype(np.uint16), count16=cnt.astype(np.uint16))
        if divisive:
            extra["divisive_columns"] = np.array([weight_name])
        if rows_nd and rows_nd > 0:
            extra["rows_nd"] = np.int64(rows_nd)
            for i, c in enumerate(chroms):
                order = np.lexsort((c.bin2, c.bin1))
                cb1 = np.asarray(c.bin1)[order]
                rp = np.searchsorted(cb1, np.arange(c.n + 1), side="left").astype(np.int64)
                extra["rows_%d" % i] = rowpack.pack_rows(rp, np.asarray(c.bin2)[order], np.asarray(c.count)[order],
                                                         c.n, rows_nd)
        with open(path, "wb") as fh:
            np.savez(fh, binsize=np.int64(binsize), chrom_names=names,
                     chrom_lengths=nb * binsize, chrom_offset=off,
                     bin1_id=b1.astype(np.int32 if off[-1] < 2**31 else np.int64),
                     bin2_id=b2.astype(np.int32 if off[-1] < 2**31 else np.int64),
                     count=cnt, **extra, **{"bins_" + weight_name: w})

    # -- per-chromosome access -------------------------------------------------
    def _cid(self, chrom: str) -> int:
        try:
            return self.chromnames.index(chrom)
        except ValueError:
            raise KeyError("chromosome %r not in %s" % (chrom, self.path)) from None

    def nbins(self, chrom: str) -> int:
        i = self._cid(chrom)
        return int(self.chrom_offset[i + 1] - self.chrom_offset[i])

    def upper_pixels(self, chrom: str):
        """(bin1, bin2, count) int32 arrays with chromosome-local bin ids."""
        i = self._cid(chrom)
        lo, hi = self._pix_range(i)
        off = self.chrom_offset[i]
        b1 = (self.bin1_id[lo:hi] - off).astype(np.int32)
        b2 = (self.bin2_id[lo:hi] - off).astype(np.int32)
        return b1, b2, np.ascontiguousarray(self.count[lo:hi], dtype=np.int32)

    def upper_pixels_csr(self, chrom: str):
        """(bin1_offset int64[n+1] rebased to 0, bin2 int32, count int32): cooler's
        ``indexes/bin1_offset`` restricted to the chromosome plus its pixel columns."""
        i = self._cid(chrom)
        lo, hi = self._pix_range(i)
        off = self.chrom_offset[i]
        n = self.nbins(chrom)
        rp = np.searchsorted(self.bin1_id[lo:hi], np.arange(off, off + n + 1), side="left").astype(np.int64)
        b2 = (self.bin2_id[lo:hi] - off).astype(np.int32)
        return rp, b2, np.ascontiguousarray(self.count[lo:hi], dtype=np.int32)

    def upper_pixels_csr16(self, chrom: str):
        """(bin1_offset int64[n+1], bin2 - bin1 uint16, count uint16), or None when the
        container has no narrow columns."""
        if self.delta16 is None or self.count16 is None:
            return None
        i = self._cid(chrom)
        lo, hi = self._pix_range(i)
        off = self.chrom_offset[i]
        n = self.nbins(chrom)
        rp = np.searchsorted(self.bin1_id[lo:hi], np.arange(off, off + n + 1), side="left").astype(np.int64)
        return rp, np.ascontiguousarray(self.delta16[lo:hi]), np.ascontiguousarray(self.count16[lo:hi])

    def weights_divisive(self, name: str) -> bool:
        return name in self.divisive

    def upper_pixels_rows(self, chrom: str, nd_min: int):
        """Packed pixel rows of the chromosome (uint8 blob for ``pk_chrom_upload_rows``) when the
        container holds them and they cover ``nd_min`` distances, else None."""
        if self.rows_nd < nd_min:
            return None
        key = self._row_keys.get(self._cid(chrom))
        return None if key is None else np.asarray(self._col(key))

    def weights(self, chrom: str, name: str) -> np.ndarray:
        if name not in self.weight_columns:
            raise KeyError("no weight column %r in %s" % (name, self.path))
        i = self._cid(chrom)
        return np.ascontiguousarray(
            self.weight_columns[name][self.chrom_offset[i]:self.chrom_offset[i + 1]],
            dtype=np.float64)


# ---------------------------------------------------------------------------
# cooler stand-in: the four calls the reference makes
# ---------------------------------------------------------------------------
class _MatrixSelector:
    def __init__(self, store: PKCool, balance, sparse: bool):
        self._s, self._balance, self._sparse = store, balance, sparse

    def fetch(self, chrom: str):
        from scipy import sparse as sp
        b1, b2, cnt = self._s.upper_pixels(chrom)
        n = self._s.nbins(chrom)
        off = b1 != b2
        row = np.concatenate([b1, b2[off]])
        col = np.concatenate([b2, b1[off]])
        data = np.concatenate([cnt, cnt[off]])
        if self._balance:
            name = "weight" if self._balance is True else self._balance
            w = self._s.weights(chrom, name)
            if self._s.weights_divisive(name):             # cooler: bias = 1 / bias for divisive_weights columns
                with np.errstate(divide="ignore", invalid="ignore"):
                    w = 1.0 / w
            data = w[row] * w[col] * data
        mat = sp.coo_matrix((data, (row, col)), shape=(n, n))
        return mat if self._sparse else mat.toarray()


class _BinsSelector:
    def __init__(self, store: PKCool):
        self._s = store

    def fetch(self, chrom: str):
        import pandas as pd
        i = self._s._cid(chrom)
        n = self._s.nbins(chrom)
        start = np.arange(n, dtype=np.int64) * self._s.binsize
        end = np.minimum(start + self._s.binsize, self._s.chrom_lengths[i])
        cols = {"chrom": [chrom] * n, "start": start, "end": end}
        for k in self._s.weight_columns:
            cols[k] = self._s.weights(chrom, k)
        return pd.DataFrame(cols)


class Cooler:
    """Stand-in for ``cooler.Cooler`` limited to what the scoring path calls."""

    def __init__(self, uri: str):
        self._s = PKCool(uri.split("::")[0])
        self.uri = uri

    @property
    def chromnames(self):
        return list(self._s.chromnames)

    @property
    def chromsizes(self):
        import pandas as pd
        return pd.Series(self._s.chrom_lengths, index=self._s.chromnames, name="length")

    @property
    def binsize(self):
        return self._s.binsize

    def matrix(self, balance=True, sparse=False, **_):
        return _MatrixSelector(self._s, balance, sparse)

    def bins(self):
        return _BinsSelector(self._s)


# ---------------------------------------------------------------------------
# product-side reader
# ---------------------------------------------------------------------------
class _RealCoolAdapter:
    """Same surface as PKCool over the real ``cooler`` package (if installed)."""

    def __init__(self, uri: str):
        import cooler  # noqa: F401 -- only reached when the package exists
        self._c = cooler.Cooler(uri)
        self.chromnames = list(self._c.chromnames)
        self.binsize = int(self._c.binsize)
        self.chrom_lengths = np.asarray(self._c.chromsizes.values, dtype=np.int64)     # `depth` (calculate_depth.py:46)

    def nbins(self, chrom):
        lo, hi = self._c.extent(chrom)
        return int(hi - lo)

    def upper_pixels(self, chrom):
        lo, _ = self._c.extent(chrom)
        df = self._c.matrix(balance=False, as_pixels=True, join=False).fetch(chrom)
        return ((df["bin1_id"].values - lo).astype(np.int32),
                (df["bin2_id"].values - lo).astype(np.int32),
                df["count"].values.astype(np.int32))

    def upper_pixels_csr(self, chrom):
        b1, b2, cnt = self.upper_pixels(chrom)
        rp = np.searchsorted(b1, np.arange(self.nbins(chrom) + 1)).astype(np.int64)
        return rp, b2, cnt

    def weights_divisive(self, name):
        try:
            return bool(self._c.open("r")["bins"][name].attrs.get("divisive_weights", False))
        except Exception:
            return False

    def weights(self, chrom, name):
        return np.ascontiguousarray(self._c.bins().fetch(chrom)[name].values, dtype=np.float64)


class H5Cool:
    """Same surface as PKCool over a real cooler file (``.cool`` / ``.mcool::/resolutions/N``),
    read with the built-in HDF5 subset reader (``h5mini``) -- no h5py / cooler needed.

    cooler's schema: ``pixels/{bin1_id,bin2_id,count}`` sorted by (bin1, bin2), upper triangle
    (``storage-mode = symmetric-upper``), genome-wide bin ids, inter-chromosomal pixels included;
    ``indexes/bin1_offset`` is the CSR row pointer of that table and ``indexes/chrom_offset`` the
    first bin of each chromosome. ``matrix(...).fetch(chrom)`` of the reference
    (``score_chromosome.py:42-43``) is the intra-chromosomal block: pixels of the chromosome's
    rows whose ``bin2`` also lies in the chromosome."""

    packs_rows_on_the_fly = True       # upper_pixels_rows packs per call and takes `scoring_weights`

    def __init__(self, uri: str):
        from . import h5mini
        path, _, group = uri.partition("::")
        self.path = path
        self._f = h5mini.File(path)
        g = self._f[group] if group.strip("/") else self._f.root
        for need in ("chroms/name", "indexes/chrom_offset", "indexes/bin1_offset", "pixels/bin2_id", "pixels/count"):
            if need not in g:
                raise KeyError("%s: not a cooler (no %s); multi-resolution files need "
                               "'file.mcool::/resolutions/<binsize>'" % (uri, need))
        self._g = g
        self.chromnames = [s.split(b"\0")[0].decode() for s in g["chroms/name"].read()]
        self.chrom_offset = g["indexes/chrom_offset"].read().astype(np.int64)
        self.chrom_lengths = g["chroms/length"].read().astype(np.int64) if "chroms/length" in g else None
        bs = g.attrs.get("bin-size")
        if bs is None and "bins/start" in g:
            se = g["bins/start"].read(0, 1), g["bins/end"].read(0, 1)
            bs = int(se[1][0] - se[0][0])
        self.binsize = int(bs) if bs is not None else None
        self._bin1_offset = g["indexes/bin1_offset"]
        self._bin2 = g["pixels/bin2_id"]
        self._count = g["pixels/count"]
        self._cache = (None, None)
        self._raw = (None, None)

    def close(self):
        self._f.close()

    def _cid(self, chrom: str) -> int:
        try:
            return self.chromnames.index(chrom)
        except ValueError:
            raise KeyError("chromosome %r not in %s" % (chrom, self.path)) from None

    def nbins(self, chrom: str) -> int:
        i = self._cid(chrom)
        return int(self.chrom_offset[i + 1] - self.chrom_offset[i])

    def _fetch(self, chrom: str):
        """(row pointer int64[n+1], bin2 int32 local, count int32) of the intra-chromosomal block."""
        if self._cache[0] == chrom:
            return self._cache[1]
        i = self._cid(chrom)
        lo, hi = int(self.chrom_offset[i]), int(self.chrom_offset[i + 1])
        n = hi - lo
        rp = self._bin1_offset.read(lo, hi + 1).astype(np.int64)
        p0 = int(rp[0])
        rp -= p0
        if rp.size != n + 1 or np.any(np.diff(rp) < 0):
            raise ValueError("%s: indexes/bin1_offset is not a row pointer of the pixel table" % self.path)
        # The rows of a chromosome also hold its inter-chromosomal pixels (several GB per chromosome in a
        # genome-wide 10 kb file): read bin2 / count in blocks of rows and keep the intra-chromosomal
        # pixels of each block only.
        BLOCK = 1 << 24                                       # pixels per block
        raw = self._raw[1] if self._raw[0] == chrom else None
        kept_b2, kept_cnt, row_counts = [], [], np.zeros(n, dtype=np.int64)
        r0 = 0
        while r0 < n:
            r1 = int(np.searchsorted(rp, rp[r0] + BLOCK, side="right")) - 1
            r1 = min(n, max(r1, r0 + 1))
            a, b = p0 + int(rp[r0]), p0 + int(rp[r1])
            if raw is not None:                              # already decoded for the packed rows
                b2, cnt = raw[1][a - p0:b - p0], raw[2][a - p0:b - p0]
            else:
                b2 = self._bin2.read(a, b)
                cnt = self._count.read(a, b)
            if cnt.dtype.kind == "f" and not np.all(cnt == np.rint(cnt)):
                raise ValueError("%s: non-integer pixel counts; the Poisson filter "
                                 "(scoreUtils.py:59-60) needs raw counts" % self.path)
            if cnt.size and (cnt.min() < 0 or cnt.max() > np.iinfo(np.int32).max):
                raise ValueError("%s: pixel counts outside int32" % self.path)
            # per-row work goes through the row pointer (reduceat over the non-empty rows), not through a
            # row index per pixel: the columns of a chr1-scale chromosome are 7 M pixels
            starts = (rp[r0:r1] - rp[r0]).astype(np.int64)
            lens = np.diff(rp[r0:r1 + 1])
            full = np.flatnonzero(lens > 0)
            if b2.size:
                # cooler stores pixels sorted by (bin1, bin2): the smallest bin2 of a row must not lie below the diagonal
                if np.any(np.minimum.reduceat(b2, starts[full]) - lo < full + r0):
                    raise ValueError("%s: pixels below the diagonal (storage-mode is not symmetric-upper)" % self.path)
            cis = b2 < hi
            if cis.all():
                row_counts[r0:r1] = lens
            else:                                            # drop inter-chromosomal pixels
                kept = np.zeros(r1 - r0, dtype=np.int64)
                kept[full] = np.add.reduceat(cis.astype(np.int64), starts[full])
                row_counts[r0:r1] = kept
                b2, cnt = b2[cis], cnt[cis]
            kept_b2.append((b2 - lo).astype(np.int32))
            kept_cnt.append(cnt.astype(np.int32, copy=False))
            r0 = r1
        rp = np.concatenate([[0], np.cumsum(row_counts)]).astype(np.int64)
        b2l = np.concatenate(kept_b2) if kept_b2 else np.zeros(0, np.int32)
        cnt = np.concatenate(kept_cnt) if kept_cnt else np.zeros(0, np.int32)
        out = (rp, np.ascontiguousarray(b2l), np.ascontiguousarray(cnt))
        self._cache = (chrom, out)
        return out

    def upper_pixels_csr(self, chrom: str):
        return self._fetch(chrom)

    # pixels (cis + trans) of one chromosome's rows up to which the raw columns are held whole for the native packer
    # (12 bytes per pixel as stored + the blob): 2^27 at least, more when the host has the memory for it
    RAW_LIMIT = 1 << 27
    try:
        import psutil as _psutil
        RAW_LIMIT = int(min(1 << 30, max(1 << 27, _psutil.virtual_memory().available // 64)))
        del _psutil
    except Exception:                                        # pragma: no cover - psutil is optional
        pass

    def _raw_columns(self, chrom: str):
        """(row pointer int64[n+1] rebased to 0, bin2 as stored (genome-wide ids), count as stored, first bin) of the
        chromosome's rows, decoded once and kept for the next call; None when the rows hold more than RAW_LIMIT
        pixels (a genome-wide file with deep inter-chromosomal rows: the block-wise ``_fetch`` bounds the memory)."""
        if self._raw[0] == chrom:
            return self._raw[1]
        i = self._cid(chrom)
        lo, hi = int(self.chrom_offset[i]), int(self.chrom_offset[i + 1])
        rp = self._bin1_offset.read(lo, hi + 1).astype(np.int64)
        p0 = int(rp[0]) if rp.size else 0
        rp -= p0
        if rp.size != hi - lo + 1 or np.any(np.diff(rp) < 0):
            raise ValueError("%s: indexes/bin1_offset is not a row pointer of the pixel table" % self.path)
        out = None
        if int(rp[-1]) <= self.RAW_LIMIT:
            out = (rp, self._bin2.read(p0, p0 + int(rp[-1])), self._count.read(p0, p0 + int(rp[-1])), lo)
        self._raw = (chrom, out)
        return out

    def prefetch(self, chrom: str) -> None:
        """Decode and keep the chromosome's pixel columns (the next upper_pixels* call for it starts from memory)."""
        if self._raw_columns(chrom) is None:
            self._fetch(chrom)

    def upper_pixels_rows(self, chrom: str, nd_min: int, scoring_weights=False):
        """Packed pixel rows of the chromosome covering ``nd_min`` distances (uint8 blob for ``pk_chrom_upload_rows``),
        packed by the library straight from the file's columns (``pk_rows_pack``: trans pixels dropped, the
        symmetric-upper / order / count checks of ``_fetch`` made on the way); None when the rows are too many to
        hold whole -- the caller then takes the block-wise columns.

        ``scoring_weights``: the scoring path passes the weights the values are balanced with (None for raw
        counts); the pixels beyond ``nd_min`` -- most of a deep genome-wide map, none of which enters the band -- are
        then reduced to those that make a bin ``valid`` which no nearer pixel does (``utils.py:146-156``). The
        default (False) keeps them all, as ``depth`` needs."""
        from . import rowpack
        raw = self._raw_columns(chrom)
        if raw is None:
            return None
        rp, b2, cnt, lo = raw
        if nd_min < 1:                                       # any coverage will do (`depth`): PKCool.write's default
            nd_min = 352
        slim = scoring_weights is not False
        try:
            return rowpack.pack_rows_native(rp, b2, cnt, rp.size - 1, int(nd_min), bin2_base=lo, valid_only_far=slim,
                                            weights=scoring_weights if slim else None)
        except ValueError as e:
            raise ValueError("%s: %s" % (self.path, e)) from None

    def upper_pixels(self, chrom: str):
        rp, b2, cnt = self._fetch(chrom)
        b1 = np.repeat(np.arange(rp.size - 1, dtype=np.int32), np.diff(rp))
        return b1, b2, cnt

    def upper_pixels_csr16(self, chrom: str):
        rp, b2, cnt = self._fetch(chrom)
        if not b2.size:
            return None
        delta = b2 - np.repeat(np.arange(rp.size - 1, dtype=np.int32), np.diff(rp))
        if int(delta.max()) > 65535 or int(delta.min()) < 0 or int(cnt.max()) > 65535:
            return None
        return rp, delta.astype(np.uint16), cnt.astype(np.uint16)

    def weights_divisive(self, name: str) -> bool:
        """cooler's column attribute `divisive_weights` (set by hic2cool for KR / VC columns): the balanced
        value divides by the weights instead of multiplying."""
        key = "bins/" + name
        if key not in self._g:
            raise KeyError("no weight column %r in %s" % (name, self.path))
        v = self._g[key].attrs.get("divisive_weights")
        try:
            return bool(np.asarray(v).ravel()[0]) if v is not None else False
        except (IndexError, ValueError):
            return False

    def weights(self, chrom: str, name: str) -> np.ndarray:
        key = "bins/" + name
        if key not in self._g:
            raise KeyError("no weight column %r in %s" % (name, self.path))
        i = self._cid(chrom)
        return np.ascontiguousarray(self._g[key].read(int(self.chrom_offset[i]), int(self.chrom_offset[i + 1])),
                                    dtype=np.float64)


def _is_hdf5(path: str) -> bool:
    from .h5mini import SIGNATURE
    try:
        with open(path, "rb") as fh:
            pos = 0
            while True:
                fh.seek(pos)
                head = fh.read(8)
                if len(head) < 8:
                    return False
                if head == SIGNATURE:
                    return True
                pos = 512 if pos == 0 else pos * 2
    except OSError:
        return False


def open_map(uri: str):
    """Open a contact map for the CUDA path: a ``.pkcool`` container or a real cooler file
    (``.cool``, ``.mcool::/resolutions/N``). Cooler files are read by the built-in HDF5 subset
    reader; a file that uses HDF5 features outside that subset falls back to the ``cooler``
    package when it is importable."""
    path = uri.split("::")[0]
    if path.endswith(".pkcool") or path.endswith(".npz"):
        return PKCool(path)
    try:
        import cooler
    except ImportError:
        cooler = None
    if cooler is not None and getattr(cooler, "Cooler", None) is Cooler:   # stand-in injected by a test
        return PKCool(path)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if _is_hdf5(path):
        from .h5mini import H5Unsupported
        try:
            return H5Cool(uri)
        except H5Unsupported as e:
            if cooler is None:
                raise RuntimeError("%s: %s is outside the built-in HDF5 reader's subset and the `cooler` "
                                   "package is not installed" % (uri, e)) from e
            return _RealCoolAdapter(uri)
    raise RuntimeError("%s is neither a .pkcool container nor an HDF5 (.cool) file" % uri)
