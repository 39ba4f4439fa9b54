"""Seeded synthetic Hi-C maps for tests and benchmarks.

The reference ships no data that fits this container, so every parity and
throughput figure is taken on maps built here: power-law distance decay,
planted loop dots, Poisson counts, balancing weights with a fraction of masked
(NaN) bins, and a sprinkle of far-off-diagonal pixels so the reference's band
trim (``scoreUtils.py:30-33``) and its ``valid_cols`` scan over *all* pixels
(``utils.py:151-156``) are both exercised.

Only the upper triangle (bin1 <= bin2) is stored, sorted by (bin1, bin2), the
way a cooler file stores pixels.
"""
from __future__ import annotations

import dataclasses
import hashlib

import numpy as np

# hg19 chromosome lengths (bp), chr1..chr22, chrX -- used for the C3/C4 shapes.
HG19_LENGTHS = {
    "chr1": 249250621, "chr2": 243199373, "chr3": 198022430, "chr4": 191154276,
    "chr5": 180915260, "chr6": 171115067, "chr7": 159138663, "chr8": 146364022,
    "chr9": 141213431, "chr10": 135534747, "chr11": 135006516, "chr12": 133851895,
    "chr13": 115169878, "chr14": 107349540, "chr15": 102531392, "chr16": 90354753,
    "chr17": 81195210, "chr18": 78077248, "chr19": 59128983, "chr20": 63025520,
    "chr21": 48129895, "chr22": 51304566, "chrX": 155270560,
}


@dataclasses.dataclass
class SynthChrom:
    """One chromosome of a synthetic map (upper triangle, cooler pixel order)."""

    name: str
    n: int                 # number of bins
    bin1: np.ndarray       # int32[nnz]
    bin2: np.ndarray       # int32[nnz]
    count: np.ndarray      # int32[nnz]
    weights: np.ndarray    # float64[n], NaN = masked bin
    loops: np.ndarray      # int64[n_loops, 2] planted (x, y) centres, x < y

    def checksum(self) -> str:
        h = hashlib.sha256()
        for a in (self.bin1, self.bin2, self.count, self.weights):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


def make_chromosome(name: str, n: int, *, seed: int, depth: float = 300.0,
                    alpha: float = 1.0, band: int = 330, n_loops: int | None = None,
                    loop_min: int = 8, loop_max: int = 250, nan_frac: float = 0.01,
                    far_pixels: int | None = None) -> SynthChrom:
    """Build one chromosome.

    lambda(d) = depth / (1 + d)**alpha for d < band; planted loops multiply a
    3x3 patch (centre x5, ring x3); counts ~ Poisson(lambda); zeros dropped.
    """
    rng = np.random.default_rng(seed)
    band = int(min(band, n))
    if n_loops is None:
        n_loops = max(4, n // 40)
    loop_max = min(loop_max, band - 3, n - 3)
    loop_min = min(loop_min, loop_max)

    # planted loops
    la = rng.integers(2, max(3, n - loop_max - 2), size=n_loops)
    ll = rng.integers(loop_min, loop_max + 1, size=n_loops)
    loops = np.unique(np.stack([la, la + ll], axis=1), axis=0)
    loops = loops[loops[:, 1] < n - 2]

    # dense band in diagonal-major layout: lam[d, x] for pixel (x, x+d)
    d = np.arange(band, dtype=np.float64)
    lam = np.repeat((depth / (1.0 + d) ** alpha)[:, None], n, axis=1)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            x = loops[:, 0] + dx
            y = loops[:, 1] + dy
            dd = y - x
            ok = (x >= 0) & (y < n) & (dd >= 0) & (dd < band)
            lam[dd[ok], x[ok]] *= 5.0 if (dx == 0 and dy == 0) else 3.0
    cnt = rng.poisson(lam).astype(np.int32)
    # slots past the end of each diagonal do not exist
    xs = np.arange(n)
    cnt[(xs[None, :] + np.arange(band)[:, None]) >= n] = 0

    dd, xx = np.nonzero(cnt)
    b1 = xx.astype(np.int32)
    b2 = (xx + dd).astype(np.int32)
    cc = cnt[dd, xx]

    # sparse far-off-diagonal pixels (beyond the band)
    if far_pixels is None:
        far_pixels = n
    if n - band > 2 and far_pixels > 0:
        fx = rng.integers(0, n - band - 1, size=far_pixels)
        fd = rng.integers(band, n, size=far_pixels)
        fy = fx + fd
        ok = fy < n
        fx, fy = fx[ok], fy[ok]
        key = np.unique(fx.astype(np.int64) * n + fy)
        b1 = np.concatenate([b1, (key // n).astype(np.int32)])
        b2 = np.concatenate([b2, (key % n).astype(np.int32)])
        cc = np.concatenate([cc, np.ones(key.size, dtype=np.int32)])

    order = np.lexsort((b2, b1))
    b1, b2, cc = b1[order], b2[order], cc[order]

    w = rng.uniform(0.7, 1.3, size=n) / np.sqrt(depth)
    n_nan = int(round(nan_frac * n))
    if n_nan:
        w[rng.choice(n, size=n_nan, replace=False)] = np.nan

    return SynthChrom(name=name, n=n, bin1=b1, bin2=b2, count=cc.astype(np.int32),
                      weights=w, loops=loops.astype(np.int64))


def make_genome(sizes: dict[str, int], *, seed: int, **kw) -> list[SynthChrom]:
    """sizes: name -> number of bins. Each chromosome gets seed + index."""
    return [make_chromosome(name, n, seed=seed + 1000 * i, **kw)
            for i, (name, n) in enumerate(sizes.items())]


def hg19_bins(res: int) -> dict[str, int]:
    return {k: -(-v // res) for k, v in HG19_LENGTHS.items()}


def band_pixels(n: int, lower: int, upper: int, w: int) -> int:
    """sum_{d=lower_eff..upper_eff} (n - d): the BASELINE.json denominator."""
    lo = max(lower, w + 1)
    up = min(upper, n - 2 * w)
    if up < lo:
        return 0
    k = up - lo + 1
    return k * n - (lo + up) * k // 2
