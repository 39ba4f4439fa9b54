"""Strict restatements of third-party arithmetic in the oracle vs the real
libraries: these are the operation orders the CUDA
kernels reproduce. CPU only."""
import numpy as np
import pytest
from scipy import stats
from scipy.ndimage import gaussian_filter
from sklearn.isotonic import IsotonicRegression

from oracle import peakachu_oracle as po


def test_gaussian_kernel_constants():
    k = po.gaussian_kernel_sigma1()
    want = [float.fromhex(h) for h in ("0x1.18a9c4fd536c6p-13", "0x1.22724cb7eb269p-8",
                                       "0x1.ba4b99d1799abp-5", "0x1.ef8eb9ad499bap-3",
                                       "0x1.9884a307594fbp-2")]
    assert k.tolist() == want


@pytest.mark.parametrize("S", [11, 13, 15])
def test_gaussian_matches_scipy_bitwise(S):
    rng = np.random.default_rng(S)
    W = rng.gamma(2.0, 1.0, size=(200, S, S)) * (rng.random((200, S, S)) > 0.2)
    G = po.gaussian_sigma1(W)
    for i in range(W.shape[0]):
        assert np.array_equal(G[i], gaussian_filter(W[i], sigma=1, order=0))


def test_pairwise_mean_matches_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [127, 128, 129, 255, 256, 1000, 1989, 24900, 24589]:
        a = rng.gamma(2.0, 1.0, size=n) * (rng.random(n) > 0.3)
        assert po.pairwise_mean(a) == a.mean(), n


def _sk_iso(d, y, maxdis):
    IR = IsotonicRegression(increasing=False, out_of_bounds="clip")
    IR.fit(d, y)
    return IR.predict(list(range(maxdis + 1)))


def test_isotonic_matches_sklearn_bitwise():
    rng = np.random.default_rng(1)
    for trial in range(400):
        maxdis = int(rng.integers(3, 700))
        base = 300.0 / (1.0 + np.arange(maxdis + 1)) ** rng.uniform(0.5, 1.5)
        noise = rng.uniform(0.0, 0.6)
        y = base * np.exp(noise * rng.standard_normal(maxdis + 1))
        if trial % 3 == 0:                       # runs of exact ties and zeros
            y = np.round(y, 1)
        if trial % 5 == 0:
            y[rng.random(maxdis + 1) < 0.2] = 0.0
        d = np.where(y > 0)[0]
        if d.size == 0:
            continue
        got = po.isotonic_nonincreasing(d, y[d], maxdis)
        want = _sk_iso(d, y[d], maxdis)
        assert np.array_equal(got, want), trial


def test_poisson_sf_close_to_scipy_and_same_decisions():
    rng = np.random.default_rng(2)
    ks = np.concatenate([rng.integers(1, 60, 3000), rng.integers(60, 5000, 500)])
    mus = np.concatenate([rng.gamma(2.0, 8.0, 3000), rng.uniform(30, 5000, 500)])
    want = stats.poisson(mus).sf(ks)
    got = np.array([po.poisson_sf(int(k), float(m)) for k, m in zip(ks, mus)])
    ok = want > 1e-290
    assert np.allclose(got[ok], want[ok], rtol=1e-10, atol=0)
    assert np.array_equal(got < 0.01, want < 0.01)
    assert po.poisson_sf(3, 0.0) == 0.0 and np.isnan(po.poisson_sf(3, float("nan")))
