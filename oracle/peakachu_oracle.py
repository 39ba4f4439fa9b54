"""CPU oracle for the Peakachu per-pixel loop-scoring path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it. The product path (``peakachu_b200``) never does
and fails loudly when its CUDA library is missing.

It restates, in numpy, the reference's algorithm for ``score_chromosome`` /
``score_genome`` (every function cites the reference file:line it follows; paths
are relative to the upstream repository root). Where the reference delegates
arithmetic to a third-party library (scipy.ndimage, scipy.stats, sklearn.isotonic,
numba's array reductions, sklearn forests) there are two flavours here:

* the ``Chromosome`` class calls the *same* library entry points the reference
  calls, except where the reference's own numba code is restated in numpy, and
* stand-alone strict restatements of that library arithmetic (``gaussian_sigma1``,
  ``pairwise_mean``, ``isotonic_nonincreasing``, ``poisson_sf``, ``forest_apply``,
  ``forest_proba``) that spell out the operation order the CUDA kernels must
  reproduce; ``tests/test_oracle_*`` check each against the real library.

Parity pin: the reference has no tests or golden vectors.
The oracle is pinned instead against outputs of the reference itself, imported
from /root/reference and run on seeded synthetic inputs; those outputs and the
script that made them are committed under ``tests/golden/``. The cooler boundary
(balanced value = (w[row]*w[col])*count on the symmetric matrix) is defined by the
stand-in in ``peakachu_b200/coolio.py`` because ``cooler`` is not installed:
**parity unpinned at the cooler boundary**.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import sparse, stats

np.seterr(divide="ignore", invalid="ignore")   # score_chromosome.py:9

BATCH = 100000   # scoreUtils.py:104


# ---------------------------------------------------------------------------
# strict restatements of third-party arithmetic (operation order spelled out)
# ---------------------------------------------------------------------------
def gaussian_kernel_sigma1() -> np.ndarray:
    """scipy.ndimage ``_gaussian_kernel1d(sigma=1, order=0, radius=4)``:
    phi = exp(-0.5 * x**2), x = -4..4, normalised by its sum (call site
    scoreUtils.py:86, ``gaussian_filter(arr, sigma=1, order=0)``; radius =
    int(4.0 * 1 + 0.5) = 4). Returned as k[0..4] = far tap .. centre tap."""
    x = np.arange(-4, 5)
    phi = np.exp(-0.5 / 1.0 * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[:5])


def _correlate1d_symmetric(X: np.ndarray, axis: int, k: np.ndarray) -> np.ndarray:
    """scipy ``correlate1d`` on its symmetric-kernel branch, mode='reflect'
    (d c b a | a b c d | d c b a):  t = x[c]*k[4]; for j=-4..-1:
    t = t + (x[c+j] + x[c-j]) * k[4+j]   -- no fused multiply-add."""
    pad = [(0, 0)] * X.ndim
    pad[axis] = (4, 4)
    P = np.pad(X, pad, mode="symmetric")
    n = X.shape[axis]

    def sl(off):
        s = [slice(None)] * X.ndim
        s[axis] = slice(4 + off, 4 + off + n)
        return P[tuple(s)]

    t = sl(0) * k[4]
    for j in (-4, -3, -2, -1):
        t = t + (sl(j) + sl(-j)) * k[4 + j]
    return t


def gaussian_sigma1(W: np.ndarray) -> np.ndarray:
    """gaussian_filter(w, sigma=1, order=0) for a stack of windows (N, S, S):
    axis 0 of each window first, then axis 1 (scipy filters the axes in order)."""
    k = gaussian_kernel_sigma1()
    t = _correlate1d_symmetric(W, 1, k)
    return _correlate1d_symmetric(t, 2, k)


def pairwise_sum(a: np.ndarray) -> float:
    """numpy's pairwise summation of a contiguous float64 vector (what
    ``ndarray.mean`` uses before dividing; utils.py:169)."""
    a = np.asarray(a, dtype=np.float64)
    n = a.size
    if n < 8:
        s = 0.0
        for v in a:
            s = s + float(v)
        return s
    if n <= 128:
        r = [float(a[i]) for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + float(a[i + j])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def pairwise_mean(a: np.ndarray) -> float:
    return pairwise_sum(a) / float(np.asarray(a).size)


def pava_nonincreasing(y: np.ndarray) -> np.ndarray:
    """scipy.optimize.isotonic_regression(y, increasing=False) with unit weights:
    reverse, pool-adjacent-violators (Busing 2022, Algorithm 1, as in scipy's
    ``_pava_pybind``), reverse back. Block means are sum/weight of running sums."""
    x = np.array(y[::-1], dtype=np.float64)
    n = x.size
    w = np.ones(n, dtype=np.float64)
    r = np.full(n + 1, -1, dtype=np.int64)
    r[0] = 0
    if n > 1:
        r[1] = 1
    b = 0
    xb_prev, wb_prev = x[0], w[0]
    i = 1
    while i < n:
        b += 1
        xb, wb = x[i], w[i]
        if xb_prev >= xb:
            b -= 1
            sb = wb_prev * xb_prev + wb * xb
            wb = wb + wb_prev
            xb = sb / wb
            while i < n - 1 and xb >= x[i + 1]:
                i += 1
                sb = sb + w[i] * x[i]
                wb = wb + w[i]
                xb = sb / wb
            while b > 0 and x[b - 1] >= xb:
                b -= 1
                sb = sb + w[b] * x[b]
                wb = wb + w[b]
                xb = sb / wb
        x[b] = xb_prev = xb
        w[b] = wb_prev = wb
        r[b + 1] = i + 1
        i += 1
    f = n - 1
    for kk in range(b, -1, -1):
        t = r[kk]
        x[t:f + 1] = x[kk]
        f = t - 1
    return x[::-1].copy()


def isotonic_nonincreasing(d: np.ndarray, y: np.ndarray, maxdis: int) -> np.ndarray:
    """``IsotonicRegression(increasing=False, out_of_bounds='clip').fit(d, y)
    .predict(range(maxdis+1))`` (utils.py:173-176): PAVA, drop interior points of
    constant runs, clip the query to [d.min, d.max], scipy ``interp1d`` linear
    which for float64 1-D data delegates to ``numpy.interp``: with knot j the last one
    <= x, the value is y[j] when x == knot j, else slope_j * (x - x_j) + y_j."""
    d = np.asarray(d, dtype=np.float64)
    yf = pava_nonincreasing(np.asarray(y, dtype=np.float64))
    if d.size == 0:
        raise ValueError("no positive expected value to fit")   # sklearn raises as well
    T = np.clip(np.arange(maxdis + 1, dtype=np.float64), d.min(), d.max())
    if d.size == 1:
        return np.repeat(yf, T.size)
    keep = np.ones(d.size, dtype=bool)
    keep[1:-1] = (yf[1:-1] != yf[:-2]) | (yf[1:-1] != yf[2:])
    xk, yk = d[keep], yf[keep]
    if xk.size == 1:
        return np.repeat(yk, T.size)
    j = np.clip(np.searchsorted(xk, T, side="right") - 1, 0, xk.size - 1)   # xk[j] <= T
    jn = np.minimum(j + 1, xk.size - 1)
    slope = (yk[jn] - yk[j]) / (xk[jn] - xk[j])
    out = slope * (T - xk[j]) + yk[j]
    exact = (xk[j] == T) | (j == xk.size - 1)
    return np.where(exact, yk[j], out)


def poisson_sf(k: int, mu: float) -> float:
    """Pr[Poisson(mu) > k] = regularised lower incomplete gamma P(k+1, mu)
    (scipy.stats.poisson.sf -> special.pdtrc; call site scoreUtils.py:59-60),
    by the textbook series / continued fraction in float64. Agrees with scipy to
    ~1e-11 relative (prefactor exp(a ln x - x - lgamma a)); used for the candidate *decision* only (p < 0.01)."""
    if not (mu >= 0.0) or math.isinf(mu):
        return float("nan") if not (mu >= 0.0) else 1.0
    if mu == 0.0:
        return 0.0
    a = float(int(math.floor(k))) + 1.0
    x = float(mu)
    lg = math.lgamma(a)
    if x < a + 1.0:
        term = 1.0 / a
        s = term
        n = a
        for _ in range(100000):
            n += 1.0
            term *= x / n
            s += term
            if term < s * 1e-17:
                break
        return s * math.exp(-x + a * math.log(x) - lg)
    # continued fraction for Q, P = 1 - Q (modified Lentz)
    tiny = 1e-300
    b = x + 1.0 - a
    c = 1.0 / tiny
    dd = 1.0 / b
    h = dd
    for i in range(1, 100000):
        an = -i * (i - a)
        b += 2.0
        dd = an * dd + b
        if abs(dd) < tiny:
            dd = tiny
        c = b + an / c
        if abs(c) < tiny:
            c = tiny
        dd = 1.0 / dd
        de = dd * c
        h *= de
        if abs(de - 1.0) < 1e-16:
            break
    return 1.0 - math.exp(-x + a * math.log(x) - lg) * h


def forest_apply(ff, X32: np.ndarray) -> np.ndarray:
    """Leaf index (tree-local node id) per sample per tree: sklearn
    ``_tree.pyx::_apply_dense``: from node 0, while left != -1:
    isnan(x) ? missing_go_to_left : (x_f32 <= threshold_f64)."""
    X32 = np.asarray(X32, dtype=np.float32)
    N = X32.shape[0]
    out = np.zeros((N, ff.n_trees), dtype=np.int32)
    rows = np.arange(N)
    for t in range(ff.n_trees):
        o = int(ff.node_offset[t])
        node = np.zeros(N, dtype=np.int64)
        active = ff.left[o + node] != -1
        while active.any():
            idx = o + node[active]
            xv = X32[rows[active], ff.feature[idx]].astype(np.float64)
            go_left = np.where(np.isnan(xv), ff.missing_left[idx] != 0, xv <= ff.threshold[idx])
            node[active] = np.where(go_left, ff.left[idx], ff.right[idx])
            active = ff.left[o + node] != -1
        out[:, t] = node
    return out


def forest_proba(ff, X32: np.ndarray) -> np.ndarray:
    """``predict_proba(X)[:, 1]``: float64 sum of leaf class-1 fractions in
    estimator order, then one divide by n_trees (sklearn _forest.py
    ``_accumulate_prediction`` + ``proba /= len(estimators_)``)."""
    leaves = forest_apply(ff, X32)
    acc = np.zeros(leaves.shape[0], dtype=np.float64)
    for t in range(ff.n_trees):
        acc = acc + ff.leaf_p1[ff.node_offset[t] + leaves[:, t]]
    return acc / float(ff.n_trees)


# ---------------------------------------------------------------------------
# the reference's own functions
# ---------------------------------------------------------------------------
def tocsr(X):
    """utils.py:10-15."""
    return sparse.csr_matrix((X.data, (X.row, X.col)), shape=X.shape, dtype=float)


def calculate_expected(M, maxdis, raw=False):
    """utils.py:139-178. Per-diagonal mean over 'valid' bins (zeros included),
    only where more than 10 entries; isotonic non-increasing fit over distances
    with a positive mean, predicted at every distance with clipping."""
    from sklearn.isotonic import IsotonicRegression
    n = M.shape[0]
    R, C = M.nonzero()
    valid_pixels = np.isfinite(M.data)
    if raw:
        R, C, data = R[valid_pixels], C[valid_pixels], M.data[valid_pixels]
        M = sparse.csr_matrix((data, (R, C)), shape=M.shape, dtype=float)
        marg = np.array(M.sum(axis=0)).ravel()
        valid_cols = marg > 0
    else:
        valid_cols = np.zeros(n, dtype=bool)
        valid_cols[R[valid_pixels]] = True       # utils.py:151-156 (set loops)
        valid_cols[C[valid_pixels]] = True
    exp_arr = np.zeros(maxdis + 1)
    for i in range(maxdis + 1):
        valid = valid_cols if i == 0 else valid_cols[:-i] * valid_cols[i:]
        diag = M.diagonal(i)
        diag = diag[valid]
        if diag.size > 10:
            exp_arr[i] = diag.mean()
    IR = IsotonicRegression(increasing=False, out_of_bounds="clip")
    _d = np.where(exp_arr > 0)[0]
    IR.fit(_d, exp_arr[_d])
    return IR.predict(list(range(maxdis + 1)))


def distance_normalize(arr_pool, exp_bychrom, xi, yi, w):
    """utils.py:211-237 with utils.py:180-202 inlined, vectorised over windows.

    Per window: NaN -> 0; reject when count_nonzero < 0.1 * size; ll_mean = numba
    ``window[:w,:w].mean()`` = sequential row-major float64 sum / (w*w); keep when
    ll_mean > 0 and centre / ll_mean > 0.1; then window / exp[|col - row|]
    (returned un-normalised if the largest distance is >= len(exp))."""
    W = np.array(arr_pool, dtype=np.float64)
    W[np.isnan(W)] = 0
    S = 2 * w + 1
    nz = np.count_nonzero(W.reshape(W.shape[0], -1), axis=1)
    ok = ~(nz < S * S * 0.1)
    s = np.zeros(W.shape[0])
    for a in range(w):
        for b in range(w):
            s = s + W[:, a, b]
    ll_mean = s / float(w * w)
    ok &= ll_mean > 0
    p2ll = W[:, w, w] / ll_mean
    ok &= p2ll > 0.1
    W, xk, yk = W[ok], xi[ok], yi[ok]
    a = np.arange(S)
    D = np.abs((yk[:, None, None] - w + a[None, None, :]) - (xk[:, None, None] - w + a[None, :, None]))
    too_far = D.reshape(D.shape[0], -1).max(axis=1) >= exp_bychrom.size if D.shape[0] else np.zeros(0, bool)
    E = exp_bychrom[np.minimum(D, exp_bychrom.size - 1)]
    normed = W / E
    if too_far.any():
        normed[too_far] = W[too_far]
    return normed, np.stack([xk, yk], axis=1)


def image_normalize(G):
    """utils.py:204-209 on a stack: (a - min) / (max - min); NaN propagates."""
    flat = G.reshape(G.shape[0], -1)
    mn = flat.min(axis=1)[:, None, None]
    mx = flat.max(axis=1)[:, None, None]
    return (G - mn) / (mx - mn)


def window_features(M, exp_arr, xi, yi, w):
    """The body of Chromosome.getwindow (scoreUtils.py:75-93) /
    trainUtils.buildmatrix (trainUtils.py:31-42) after the coordinate mask."""
    S = 2 * w + 1
    if xi.size == 0:
        return np.zeros((0, S * S)), np.zeros((0, 2), dtype=np.int64)
    seed = np.arange(-w, w + 1)
    delta = np.tile(seed, (seed.size, 1))
    xxx = xi.reshape((xi.size, 1, 1)) + delta.T
    yyy = yi.reshape((yi.size, 1, 1)) + delta
    v = np.array(M[xxx.ravel(), yyy.ravel()]).ravel()
    vvv = v.reshape((xi.size, S, S))
    windows, clist = distance_normalize(vvv, exp_arr, xi, yi, w)
    if windows.shape[0] == 0:
        return np.zeros((0, S * S)), np.zeros((0, 2), dtype=np.int64)
    fea = image_normalize(gaussian_sigma1(windows)).reshape(windows.shape[0], S * S)
    return fea, clist


def buildmatrix(Matrix, coords, w=5):
    """trainUtils.py:12-44 (training-set features; matrix is NOT NaN-trimmed)."""
    coords = np.r_[coords]
    xi, yi = coords[:, 0], coords[:, 1]
    mask = (xi - w >= 0) & (yi + w + 1 <= Matrix.shape[0]) & (yi - xi > w)
    xi, yi = xi[mask], yi[mask]
    if xi.size < 10:
        return None
    maxdis = int(np.abs(xi - yi).max()) + 2 * w
    exp_arr = calculate_expected(Matrix, maxdis)
    fea, _ = window_features(Matrix, exp_arr, xi, yi, w)
    return list(fea)


class Chromosome:
    """scoreUtils.py:9-135, same constructor, attributes and methods."""

    def __init__(self, M, model, raw_M=None, weights=None, lower=6, upper=300,
                 cname="chrm", res=10000, width=5):
        lower = max(lower, width + 1)                       # :13
        upper = min(upper, M.shape[0] - 2 * width)          # :14
        if weights is None:                                 # :16-24
            self.exp_arr = calculate_expected(M, upper + 2 * width, raw=True)
            if M is raw_M:
                self.background = self.exp_arr
            else:
                self.background = calculate_expected(raw_M, upper + 2 * width, raw=True)
        else:
            self.exp_arr = calculate_expected(M, upper + 2 * width, raw=False)
            self.background = self.exp_arr
        self.raw_M = raw_M
        self.weights = weights
        R, C = M.nonzero()                                  # :30-33
        validmask = np.isfinite(M.data) & (C - R > (-2 * width)) & (C - R < (upper + 2 * width))
        R, C, data = R[validmask], C[validmask], M.data[validmask]
        self.M = sparse.csr_matrix((data, (R, C)), shape=M.shape)
        self.get_candidate(lower, upper)
        self.chromname = cname
        self.r = res
        self.w = width
        self.model = model
        self.lower, self.upper = lower, upper

    def get_candidate(self, lower, upper):
        """scoreUtils.py:40-68: raw count > 0, Poisson upper tail finite and < 0.01
        against expected-raw = background[d] / (w_x * w_y). Order: d asc, x asc."""
        xs, ys, ps = [], [], []
        idx = np.arange(self.raw_M.shape[0])
        for i in range(lower, upper + 1):
            diag = self.raw_M.diagonal(i)
            e = self.background[i]
            if (diag.size > 0) and (e > 0):
                xi, yi = idx[:-i], idx[i:]
                if self.weights is None:
                    exp = np.ones(diag.size, dtype=float) * e
                else:
                    exp = np.ones(diag.size, dtype=float) * e / (self.weights[:-i] * self.weights[i:])
                pvalues = stats.poisson(exp).sf(diag)
                mask = (diag > 0) & np.isfinite(pvalues)
                xs.append(xi[mask]); ys.append(yi[mask]); ps.append(pvalues[mask])
        x_arr = np.concatenate(xs) if xs else np.array([], dtype=int)
        y_arr = np.concatenate(ys) if ys else np.array([], dtype=int)
        p_arr = np.concatenate(ps) if ps else np.array([], dtype=float)
        mask = p_arr < 0.01
        self.ridx, self.cidx = x_arr[mask], y_arr[mask]

    def getwindow(self, coords):
        """scoreUtils.py:70-93."""
        w = self.w
        coords = np.asarray(coords).reshape(-1, 2)
        xi, yi = coords[:, 0], coords[:, 1]
        mask = (xi - w >= 0) & (yi + w + 1 <= self.M.shape[0])
        return window_features(self.M, self.exp_arr, xi[mask], yi[mask], w)

    def score(self, thre=0.5):
        """scoreUtils.py:95-125, including the 100,000-candidate batches and the
        silent drop of a batch with <= 1 surviving window (:108)."""
        print("scoring matrix {}".format(self.chromname))
        print("number of candidates {}".format(self.ridx.size))
        ri, ci, pp = [], [], []
        for t in range(0, self.ridx.size, BATCH):
            coords = np.stack([self.ridx[t:t + BATCH], self.cidx[t:t + BATCH]], axis=1)
            fea, clist = self.getwindow(coords)
            if fea.shape[0] > 1:
                p = self.model.predict_proba(fea)[:, 1]
                pf = p > thre
                ri.append(clist[:, 0][pf]); ci.append(clist[:, 1][pf]); pp.append(p[pf])
        ri = np.concatenate(ri).astype(int) if ri else np.zeros(0, dtype=int)
        ci = np.concatenate(ci).astype(int) if ci else np.zeros(0, dtype=int)
        pp = np.concatenate(pp) if pp else np.zeros(0)
        result = sparse.csr_matrix((pp, (ri, ci)), shape=self.M.shape)
        if ri.size > 0:
            data = np.array(self.M[ri, ci]).ravel()
            self.M = sparse.csr_matrix((data, (ri, ci)), shape=self.M.shape)
        else:
            self.M = result
        return result, self.M

    def writeBed(self, outfil, prob_csr, raw_csr):
        """scoreUtils.py:127-135: append; rows in CSR (x asc, y asc) order; floats
        through str(numpy.float64)."""
        r, c = prob_csr.nonzero()
        pv = np.asarray(prob_csr[r, c]).ravel() if r.size else np.zeros(0)
        rv = np.asarray(raw_csr[r, c]).ravel() if r.size else np.zeros(0)
        with open(outfil, "a") as out:
            for i in range(r.size):
                line = [self.chromname, r[i] * self.r, (r[i] + 1) * self.r,
                        self.chromname, c[i] * self.r, (c[i] + 1) * self.r, pv[i], rv[i]]
                out.write("\t".join(list(map(str, line))) + "\n")


def score_map(lib, model, chrom_names, *, weight_name="weight", lower=6, upper=300,
              res=10000, min_prob=0.5, output=None, genome=False):
    """Body of score_chromosome.main / score_genome.main (score_chromosome.py:14-71,
    score_genome.py:46-84) on an already-open cooler-like ``lib``. The two drivers label the
    records differently: score_chromosome.py:37-38 strips leading 'c', 'h', 'r' characters and
    prepends 'chr'; score_genome.py:48-51 (``genome=True``) prepends 'chr' unless the name starts with it."""
    import io
    from contextlib import redirect_stdout
    width = int((np.sqrt(model.feature_importances_.size) - 1) / 2)
    correct = False if weight_name.lower() == "raw" else weight_name
    stats_out = []
    for key in chrom_names:
        if genome:
            cname = key if key.startswith("chr") else "chr" + key      # score_genome.py:48-51
        else:
            cname = "chr" + key.lstrip("chr")                          # score_chromosome.py:37-38
        if correct:
            M = tocsr(lib.matrix(balance=correct, sparse=True).fetch(key))
            raw_M = tocsr(lib.matrix(balance=False, sparse=True).fetch(key))
            weights = lib.bins().fetch(key)[correct].values
            X = Chromosome(M, model=model, raw_M=raw_M, weights=weights, cname=cname,
                           lower=lower, upper=upper, res=res, width=width)
        else:
            M = tocsr(lib.matrix(balance=False, sparse=True).fetch(key))
            X = Chromosome(M, model=model, raw_M=M, weights=None, cname=cname,
                           lower=lower, upper=upper, res=res, width=width)
        with redirect_stdout(io.StringIO()):
            result, R = X.score(thre=min_prob)
        if output is not None:
            X.writeBed(output, result, R)
        stats_out.append(dict(chrom=key, n=M.shape[0], candidates=int(X.ridx.size), rows=int(result.nnz)))
    return stats_out
