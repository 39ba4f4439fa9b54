"""Contact-map ingestion for the scoring path.

The reference reads a ``.cool`` through four ``cooler`` calls
(``score_chromosome.py:33-44``, ``score_genome.py:28-61``)::

    Lib = cooler.Cooler(uri)
    Lib.chromnames
    Lib.matrix(balance=<name>|False, sparse=True).fetch(chrom)   # full symmetric COO
    Lib.bins().fetch(chrom)[<name>].values                       # float64[n]

``cooler``/``h5py`` are not installed in this image, so this module provides

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
consumes), from a ``.pkcool`` or -- when the real ``cooler`` package exists --
from a real ``.cool`` URI.
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
        self.binsize = int(z["binsize"])
        self.chromnames = [str(s) for s in z["chrom_names"]]
        self.chrom_lengths = z["chrom_lengths"].astype(np.int64)      # bp
        self.chrom_offset = z["chrom_offset"].astype(np.int64)        # bins, len nchrom+1
        self.bin1_id = z["bin1_id"]
        self.bin2_id = z["bin2_id"]
        self.count = z["count"]
        # optional narrow pixel columns (bin2 - bin1 and count as uint16), written when every
        # pixel is representable: half the bytes to move to the GPU
        self.delta16 = z["delta16"] if "delta16" in z.files else None
        self.count16 = z["count16"] if "count16" in z.files else None
        self.weight_columns = {k[len("bins_"):]: z[k] for k in z.files if k.startswith("bins_")}
        # pixel range of each chromosome (bin1 sorted, intra-chromosomal only)
        self._pix_lo = np.searchsorted(self.bin1_id, self.chrom_offset[:-1], side="left")
        self._pix_hi = np.searchsorted(self.bin1_id, self.chrom_offset[1:], side="left")

    # -- writer -----------------------------------------------------------
    @staticmethod
    def write(path: str, chroms, binsize: int, weight_name: str = "weight") -> None:
        """chroms: iterable of synth.SynthChrom-like objects
        (name, n, bin1, bin2, count, weights)."""
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
        lo, hi = self._pix_lo[i], self._pix_hi[i]
        off = self.chrom_offset[i]
        b1 = (self.bin1_id[lo:hi] - off).astype(np.int32)
        b2 = (self.bin2_id[lo:hi] - off).astype(np.int32)
        return b1, b2, np.ascontiguousarray(self.count[lo:hi], dtype=np.int32)

    def upper_pixels_csr(self, chrom: str):
        """(bin1_offset int64[n+1] rebased to 0, bin2 int32, count int32): cooler's
        ``indexes/bin1_offset`` restricted to the chromosome plus its pixel columns."""
        i = self._cid(chrom)
        lo, hi = self._pix_lo[i], self._pix_hi[i]
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
        lo, hi = self._pix_lo[i], self._pix_hi[i]
        off = self.chrom_offset[i]
        n = self.nbins(chrom)
        rp = np.searchsorted(self.bin1_id[lo:hi], np.arange(off, off + n + 1), side="left").astype(np.int64)
        return rp, np.ascontiguousarray(self.delta16[lo:hi]), np.ascontiguousarray(self.count16[lo:hi])

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

    def nbins(self, chrom):
        lo, hi = self._c.extent(chrom)
        return int(hi - lo)

    def upper_pixels(self, chrom):
        lo, _ = self._c.extent(chrom)
        df = self._c.matrix(balance=False, as_pixels=True, join=False).fetch(chrom)
        return ((df["bin1_id"].values - lo).astype(np.int32),
                (df["bin2_id"].values - lo).astype(np.int32),
                df["count"].values.astype(np.int32))

    def weights(self, chrom, name):
        return np.ascontiguousarray(self._c.bins().fetch(chrom)[name].values, dtype=np.float64)


def open_map(uri: str):
    """Open a contact map for the CUDA path: ``.pkcool`` container, or a real
    ``.cool`` URI when the ``cooler`` package is importable."""
    path = uri.split("::")[0]
    if path.endswith(".pkcool") or path.endswith(".npz"):
        return PKCool(path)
    try:
        import cooler  # noqa: F401
    except ImportError as e:
        raise RuntimeError(
            "%s is not a .pkcool container and the `cooler` package (HDF5) is not "
            "installed in this environment" % uri) from e
    if getattr(cooler, "Cooler", None) is Cooler:   # stand-in injected by a test
        return PKCool(path)
    return _RealCoolAdapter(uri)
