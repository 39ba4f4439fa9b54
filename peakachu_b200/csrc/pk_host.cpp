// Host-side arithmetic of the peakachu_b200 library: the Poisson decision table and
// the expected-curve fit. Compiled with -ffp-contract=off: every float64 operation
// below is one IEEE operation, in the order the reference's libraries perform it.
#include <algorithm>
#include <charconv>
#include <cstring>
#include <string>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <thread>
#include <vector>

#include <stdint.h>

#include "peakachu_b200.h"

void pk_set_error(const char* fmt, ...);

// ---------------------------------------------------------------------------
// Poisson decision table.
// Reference: scoreUtils.py:59-67 keeps a pixel when stats.poisson(mu).sf(k) < 0.01.
// sf(k, mu) = P(k+1, mu) (regularised lower incomplete gamma) is strictly
// increasing in mu, so the test is mu < crit[k] with crit[k] the smallest float64
// whose upper tail reaches 0.01. The table is evaluated in 80-bit long double
// (series for P(a, x), x < a; prefactor through log1p so no large terms cancel).
// ---------------------------------------------------------------------------
namespace {

long double stirling_corr(long double a) {
    long double a2 = a * a, a3 = a2 * a, a5 = a3 * a2, a7 = a5 * a2, a9 = a7 * a2;
    return 1.0L / (12.0L * a) - 1.0L / (360.0L * a3) + 1.0L / (1260.0L * a5) - 1.0L / (1680.0L * a7) +
           1.0L / (1188.0L * a9);
}

// returns P(a, mu) and, through *pref, mu^a e^-mu / Gamma(a+1)
long double lower_gamma_p(long double a, long double mu, long double* pref) {
    long double lp;
    if (a < 32.0L) {
        lp = a * logl(mu) - mu - lgammal(a + 1.0L);
    } else {
        const long double two_pi = 6.283185307179586476925286766559L;
        long double u = (mu - a) / a;
        lp = a * (log1pl(u) - u) - 0.5L * logl(two_pi * a) - stirling_corr(a);
    }
    long double term = 1.0L, s = 1.0L, den = a;
    for (int it = 0; it < 10000000; ++it) {
        den += 1.0L;
        term *= mu / den;
        s += term;
        if (term < s * 1e-22L) break;
    }
    long double p = expl(lp);
    if (pref) *pref = p;
    return s * p;
}

double critical_mu(int k) {
    const long double target = 0.01L;
    const long double a = (long double)k + 1.0L;
    long double lo = 0.0L, hi = a;                       // sf(lo) = 0 < target <= sf(hi)
    long double t = 1.0L - 1.0L / (9.0L * a) - 2.3263478740408408L * sqrtl(1.0L / (9.0L * a));
    long double mu = a * t * t * t;                      // Wilson-Hilferty start
    if (!(mu > lo && mu < hi)) mu = 0.5L * a;
    for (int it = 0; it < 200; ++it) {
        long double pref;
        long double f = lower_gamma_p(a, mu, &pref) - target;
        if (f < 0) lo = mu; else hi = mu;
        long double pmf = pref * a / mu;                 // d sf / d mu
        long double nx = mu - f / pmf;
        if (!(nx > lo && nx < hi)) nx = 0.5L * (lo + hi);
        if (fabsl(nx - mu) <= 1e-19L * mu || (double)lo == (double)hi) { mu = nx; break; }
        mu = nx;
    }
    double m = (double)mu;
    // smallest float64 m with sf(k, m) >= target
    for (int g = 0; g < 64 && lower_gamma_p(a, (long double)m, nullptr) < target; ++g) m = std::nextafter(m, INFINITY);
    for (int g = 0; g < 64; ++g) {
        double prev = std::nextafter(m, 0.0);
        if (lower_gamma_p(a, (long double)prev, nullptr) >= target) m = prev; else break;
    }
    return m;
}

std::mutex g_tab_mu;
std::vector<double> g_tab;    // index k; [0] = smallest mu with 1 - exp(-mu) >= 0.01

}  // namespace

int pk_poisson_table_host(int32_t k_max, const double** out) {
    if (k_max < 0 || k_max > (1 << 24)) {
        pk_set_error("pk_poisson_table_host: k_max %d out of range", k_max);
        return PK_EINVAL;
    }
    std::lock_guard<std::mutex> lk(g_tab_mu);
    size_t have = g_tab.size();
    if (have < (size_t)k_max + 1) {
        // grow geometrically; fill in parallel
        size_t want = std::max<size_t>((size_t)k_max + 1, std::max<size_t>(4096, have * 2));
        g_tab.resize(want);
        unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([=]() {
                for (size_t k = have + t; k < want; k += nt) g_tab[k] = critical_mu((int)k);
            });
        for (auto& x : th) x.join();
    }
    *out = g_tab.data();
    return PK_OK;
}

extern "C" int pk_poisson_critical_mu(int32_t k_max, double* out) {
    const double* tab = nullptr;
    int r = pk_poisson_table_host(k_max, &tab);
    if (r != PK_OK) return r;
    std::lock_guard<std::mutex> lk(g_tab_mu);
    std::copy(g_tab.begin(), g_tab.begin() + k_max + 1, out);
    return PK_OK;
}

// ---------------------------------------------------------------------------
// Expected curve: utils.py:160-176.
//   exp[d] = sum[d] / cnt[d] where cnt[d] > 10, else 0
//   IsotonicRegression(increasing=False, out_of_bounds='clip').fit(d : exp>0).predict(0..len-1)
// which is scipy's PAVA (Busing 2022, Alg. 1) on the reversed sequence, sklearn's
// removal of interior points of constant runs, np.clip of the query, and
// numpy.interp (y[j] when x hits knot j, else slope*(x - x[j]) + y[j]).
// ---------------------------------------------------------------------------
int pk_fit_expected_host(const double* sum, const long long* cnt, int32_t len, double* out_exp) {
    std::vector<double> xs, ys;
    for (int32_t d = 0; d < len; ++d) {
        double e = 0.0;
        if (cnt[d] > 10) e = sum[d] / (double)cnt[d];
        if (e > 0) { xs.push_back((double)d); ys.push_back(e); }
    }
    const int n = (int)xs.size();
    if (n == 0) {
        pk_set_error("expected curve: no distance has a positive mean (reference raises in IsotonicRegression.fit)");
        return PK_EINVAL;
    }
    // PAVA on the reversed values, unit weights
    std::vector<double> x(n), w(n, 1.0);
    std::vector<int> r(n + 1, -1);
    for (int i = 0; i < n; ++i) x[i] = ys[n - 1 - i];
    r[0] = 0;
    if (n > 1) r[1] = 1;
    int b = 0;
    double xb_prev = x[0], wb_prev = w[0];
    for (int i = 1; i < n; ++i) {
        b++;
        double xb = x[i], wb = w[i];
        if (xb_prev >= xb) {
            b--;
            double sb = wb_prev * xb_prev + wb * xb;
            wb = wb + wb_prev;
            xb = sb / wb;
            while (i < n - 1 && xb >= x[i + 1]) {
                i++;
                sb = sb + w[i] * x[i];
                wb = wb + w[i];
                xb = sb / wb;
            }
            while (b > 0 && x[b - 1] >= xb) {
                b--;
                sb = sb + w[b] * x[b];
                wb = wb + w[b];
                xb = sb / wb;
            }
        }
        x[b] = xb_prev = xb;
        w[b] = wb_prev = wb;
        r[b + 1] = i + 1;
    }
    int f = n - 1;
    for (int k = b; k >= 0; --k) {
        int t = r[k];
        double xk = x[k];
        for (int i = f; i >= t; --i) x[i] = xk;
        f = t - 1;
    }
    std::vector<double> yf(n);
    for (int i = 0; i < n; ++i) yf[i] = x[n - 1 - i];
    // knots
    std::vector<double> kx, ky;
    for (int i = 0; i < n; ++i) {
        bool keep = (i == 0 || i == n - 1) || (yf[i] != yf[i - 1]) || (yf[i] != yf[i + 1]);
        if (keep) { kx.push_back(xs[i]); ky.push_back(yf[i]); }
    }
    const int m = (int)kx.size();
    const double xmin = xs[0], xmax = xs[n - 1];
    for (int32_t d = 0; d < len; ++d) {
        double T = std::min(std::max((double)d, xmin), xmax);
        if (m == 1) { out_exp[d] = ky[0]; continue; }
        int j = (int)(std::upper_bound(kx.begin(), kx.end(), T) - kx.begin()) - 1;   // kx[j] <= T
        j = std::min(std::max(j, 0), m - 1);
        if (j == m - 1 || kx[j] == T) { out_exp[d] = ky[j]; continue; }
        double slope = (ky[j + 1] - ky[j]) / (kx[j + 1] - kx[j]);
        out_exp[d] = slope * (T - kx[j]) + ky[j];
    }
    return PK_OK;
}

// test hook: the fit alone, HOST arrays
extern "C" int pk_fit_expected(const double* sum, const int64_t* cnt, int32_t len, double* out_exp) {
    std::vector<long long> c(cnt, cnt + len);
    return pk_fit_expected_host(sum, c.data(), len, out_exp);
}

// ---------------------------------------------------------------------------
// bedpe text (scoreUtils.py:127-135). The reference prints floats through
// str(numpy.float64), which is the shortest round-trip representation in Python's repr
// layout: fixed notation for 1e-4 <= |x| < 1e16 (with ".0" appended to integral values),
// otherwise d.ddde+XX with at least two exponent digits.
// ---------------------------------------------------------------------------
static char* put_repr(char* p, double v) {
    if (std::isnan(v)) { memcpy(p, "nan", 3); return p + 3; }
    if (std::isinf(v)) { if (v < 0) *p++ = '-'; memcpy(p, "inf", 3); return p + 3; }
    if (std::signbit(v)) { *p++ = '-'; v = -v; }
    if (v == 0.0) { memcpy(p, "0.0", 3); return p + 3; }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);   // shortest digits
    // buf = d[.ddd]e[+-]XX
    char digits[32];
    int nd = 0;
    char* q = buf;
    while (q < res.ptr && *q != 'e') { if (*q != '.') digits[nd++] = *q; ++q; }
    int e10 = 0;
    std::from_chars(q + 1 + (q[1] == '+'), res.ptr, e10);
    const int decpt = e10 + 1;                       // value = 0.digits * 10^decpt
    if (decpt > 16 || decpt < -3) {
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, digits + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        int ex = decpt - 1;
        *p++ = ex < 0 ? '-' : '+';
        if (ex < 0) ex = -ex;
        if (ex < 10) *p++ = '0';
        p = std::to_chars(p, p + 8, ex).ptr;
        return p;
    }
    if (decpt <= 0) {
        *p++ = '0'; *p++ = '.';
        for (int i = 0; i < -decpt; ++i) *p++ = '0';
        memcpy(p, digits, nd); return p + nd;
    }
    if (nd <= decpt) {
        memcpy(p, digits, nd); p += nd;
        for (int i = nd; i < decpt; ++i) *p++ = '0';
        *p++ = '.'; *p++ = '0';
        return p;
    }
    memcpy(p, digits, decpt); p += decpt;
    *p++ = '.';
    memcpy(p, digits + decpt, nd - decpt);
    return p + (nd - decpt);
}

extern "C" int pk_format_bedpe(const char* chrom, int64_t res, const int32_t* x, const int32_t* y, const double* prob,
                               const double* val, int64_t n, char* out, int64_t capacity, int64_t* written) {
    if (!chrom || !written || n < 0 || (n > 0 && (!x || !y || !prob || !val || !out))) {
        pk_set_error("pk_format_bedpe: bad argument");
        return PK_EINVAL;
    }
    const size_t cl = strlen(chrom);
    const int64_t per_row = 2 * (int64_t)cl + 4 * 21 + 2 * 26 + 8;      // generous upper bound per line
    if (capacity < n * per_row) {
        *written = n * per_row;
        pk_set_error("pk_format_bedpe: need %lld bytes", (long long)(n * per_row));
        return PK_ECAPACITY;
    }
    char* p = out;
    for (int64_t i = 0; i < n; ++i) {
        // the reference multiplies numpy int32 bin indices by the resolution: int32 arithmetic
        const int32_t a0 = (int32_t)((int64_t)x[i] * res), a1 = (int32_t)(((int64_t)x[i] + 1) * res);
        const int32_t b0 = (int32_t)((int64_t)y[i] * res), b1 = (int32_t)(((int64_t)y[i] + 1) * res);
        memcpy(p, chrom, cl); p += cl; *p++ = '\t';
        p = std::to_chars(p, p + 12, a0).ptr; *p++ = '\t';
        p = std::to_chars(p, p + 12, a1).ptr; *p++ = '\t';
        memcpy(p, chrom, cl); p += cl; *p++ = '\t';
        p = std::to_chars(p, p + 12, b0).ptr; *p++ = '\t';
        p = std::to_chars(p, p + 12, b1).ptr; *p++ = '\t';
        p = put_repr(p, prob[i]); *p++ = '\t';
        p = put_repr(p, val[i]); *p++ = '\n';
    }
    *written = p - out;
    return PK_OK;
}

// ---------------------------------------------------------------------------
// HDF5 chunk decoding for the .cool reader (peakachu_b200/h5mini.py): the chunks of one 1-D dataset that cover
// elements [lo, hi) are inflated (zlib) and un-shuffled (HDF5 shuffle filter: byte planes -> elements) on a few
// host threads, straight from the memory-mapped file into the caller's array. A chromosome of a genome-wide
// cooler is hundreds of megabytes of pixel columns; in Python this step was 85 % of score_chromosome's wall time.
// ---------------------------------------------------------------------------
#include <zlib.h>

#include <atomic>
#include <thread>

// HDF5 shuffle filter, undone: plane j (chunk_elems bytes apart) holds byte j of every element. 8- and 4-byte
// elements go through byte-matrix transposes in registers (8 x 8 / 4 x 4: swap the off-diagonal blocks of size
// 1, 2, 4), eight / four elements per step; a byte-strided store loop was 40 % of the decode time.
static void pk_unshuffle(const uint8_t* planes, size_t pitch, int es, int64_t cnt, uint8_t* dst) {
    int64_t k = 0;
    if (es == 8) {
        for (; k + 8 <= cnt; k += 8) {
            uint64_t r[8];
            for (int j = 0; j < 8; ++j) memcpy(&r[j], planes + (size_t)j * pitch + k, 8);
            for (int m = 1; m <= 4; m <<= 1) {
                const uint64_t mask = m == 1 ? 0x00FF00FF00FF00FFull : (m == 2 ? 0x0000FFFF0000FFFFull : 0x00000000FFFFFFFFull);
                const int sh = 8 * m;
                for (int j = 0; j < 8; ++j) {
                    if (j & m) continue;
                    const uint64_t a = r[j], b = r[j + m];
                    r[j] = (a & mask) | ((b & mask) << sh);
                    r[j + m] = ((a >> sh) & mask) | (b & ~mask);
                }
            }
            memcpy(dst + (size_t)k * 8, r, 64);
        }
    } else if (es == 4) {
        for (; k + 4 <= cnt; k += 4) {
            uint32_t r[4];
            for (int j = 0; j < 4; ++j) memcpy(&r[j], planes + (size_t)j * pitch + k, 4);
            for (int m = 1; m <= 2; m <<= 1) {
                const uint32_t mask = m == 1 ? 0x00FF00FFu : 0x0000FFFFu;
                const int sh = 8 * m;
                for (int j = 0; j < 4; ++j) {
                    if (j & m) continue;
                    const uint32_t a = r[j], b = r[j + m];
                    r[j] = (a & mask) | ((b & mask) << sh);
                    r[j + m] = ((a >> sh) & mask) | (b & ~mask);
                }
            }
            memcpy(dst + (size_t)k * 4, r, 16);
        }
    } else if (es == 2) {
        for (; k < cnt; ++k) {
            const uint16_t v = (uint16_t)(planes[k] | ((uint16_t)planes[pitch + k] << 8));
            memcpy(dst + (size_t)k * 2, &v, 2);
        }
    }
    for (; k < cnt; ++k)
        for (int j = 0; j < es; ++j) dst[(size_t)k * es + j] = planes[(size_t)j * pitch + k];
}

extern "C" int pk_h5_decode_chunks(const uint8_t* file, int64_t n_chunks, const int64_t* chunk_off, const int64_t* chunk_bytes,
                                   const int64_t* first_elem, int64_t chunk_elems, int32_t elem_size, int32_t deflate,
                                   int32_t shuffle, int32_t fletcher32, int64_t lo, int64_t hi, void* out, int32_t n_threads) {
    if (!file || n_chunks < 0 || (n_chunks > 0 && (!chunk_off || !chunk_bytes || !first_elem)) || chunk_elems <= 0 ||
        elem_size <= 0 || elem_size > 16 || hi < lo || !out) {
        pk_set_error("pk_h5_decode_chunks: bad argument");
        return PK_EINVAL;
    }
    const size_t raw_bytes = (size_t)chunk_elems * (size_t)elem_size;
    std::atomic<int64_t> next(0);
    std::atomic<int> failed(0);
    auto work = [&]() {
        std::vector<uint8_t> tmp(deflate ? raw_bytes : 0);
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_chunks || failed.load()) return;
            const uint8_t* src = file + chunk_off[i];
            size_t n_src = (size_t)chunk_bytes[i];
            if (fletcher32) {                                  // the checksum trails the filtered bytes
                if (n_src < 4) { failed.store(1); return; }
                n_src -= 4;
            }
            const uint8_t* body = src;
            if (deflate) {
                uLongf got = (uLongf)raw_bytes;
                if (uncompress(tmp.data(), &got, src, (uLong)n_src) != Z_OK || (size_t)got != raw_bytes) { failed.store(2); return; }
                body = tmp.data();
            } else if (n_src < raw_bytes) { failed.store(3); return; }
            // elements of this chunk inside [lo, hi)
            const int64_t e0 = std::max(lo, first_elem[i]), e1 = std::min(hi, first_elem[i] + chunk_elems);
            if (e1 <= e0) continue;
            const int64_t k0 = e0 - first_elem[i], cnt = e1 - e0;
            uint8_t* dst = static_cast<uint8_t*>(out) + (size_t)(e0 - lo) * elem_size;
            if (!shuffle || elem_size == 1) {
                memcpy(dst, body + (size_t)k0 * elem_size, (size_t)cnt * elem_size);
            } else {
                pk_unshuffle(body + k0, (size_t)chunk_elems, elem_size, cnt, dst);
            }
        }
    };
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(nt, 32), n_chunks));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    if (failed.load()) {
        pk_set_error("pk_h5_decode_chunks: a chunk could not be decoded (%s)",
                     failed.load() == 2 ? "inflate failed or size mismatch" : "truncated chunk");
        return PK_EINVAL;
    }
    return PK_OK;
}

// ---------------------------------------------------------------------------
// Packed pixel rows straight from cooler's columns (the format of pk_chrom_upload_rows; the same bytes
// peakachu_b200/rowpack.py::pack_rows writes): one threaded pass counts every row's band pixels, escapes and far
// pixels, a serial prefix over the rows places them, a second threaded pass writes the sections. The columns are
// the rows of ONE chromosome as the file stores them: genome-wide bin2 ids (pixels whose bin2 lies behind the
// chromosome are inter-chromosomal and dropped, score_chromosome.py:42-43 fetches the cis block), duplicates summed
// and zero counts dropped as utils.tocsr would (utils.py:10-15).
// ---------------------------------------------------------------------------
namespace {

struct RowsIn {
    const int64_t* rp;      // [n + 1], relative to the first pixel handed in
    const void* b2;
    int b2_bytes;           // 4 | 8
    int64_t b2_base;        // first bin of the chromosome (subtracted)
    const void* cnt;
    int cnt_kind;           // 0: int32, 1: int64, 2: float64
    int64_t n;
    int64_t nd;
};

struct RowTally { uint32_t band; uint32_t esc; int64_t far; };

enum { ROWS_OK = 0, ROWS_LOWER = 1, ROWS_ORDER = 2, ROWS_NEG = 3, ROWS_BIG = 4, ROWS_FRAC = 5 };

inline int64_t rows_b2(const RowsIn& in, int64_t p) {
    return (in.b2_bytes == 8 ? static_cast<const int64_t*>(in.b2)[p] : (int64_t)static_cast<const int32_t*>(in.b2)[p]) - in.b2_base;
}
inline bool rows_cnt(const RowsIn& in, int64_t p, int64_t* out) {       // false: not an integer
    if (in.cnt_kind == 0) { *out = static_cast<const int32_t*>(in.cnt)[p]; return true; }
    if (in.cnt_kind == 1) { *out = static_cast<const int64_t*>(in.cnt)[p]; return true; }
    const double v = static_cast<const double*>(in.cnt)[p];
    if (!(v == std::floor(v)) || !(std::fabs(v) < 9.0e18)) return false;
    *out = (int64_t)v;
    return true;
}

// walks the de-duplicated cis pixels of row x in order: f(d, count) for count > 0
template <class F>
inline int rows_walk(const RowsIn& in, int64_t x, F&& f) {
    int64_t p = in.rp[x];
    const int64_t pe = in.rp[x + 1];
    int64_t prev = -1;
    while (p < pe) {
        const int64_t c2 = rows_b2(in, p);
        if (c2 >= in.n) {                                   // the rest of the row is inter-chromosomal (sorted by bin2)
            for (int64_t q = p + 1; q < pe; ++q)
                if (rows_b2(in, q) < in.n) return ROWS_ORDER;
            break;
        }
        if (c2 < x) return ROWS_LOWER;
        if (c2 < prev) return ROWS_ORDER;
        int64_t sum = 0;
        while (p < pe && rows_b2(in, p) == c2) {            // duplicates are summed
            int64_t v;
            if (!rows_cnt(in, p, &v)) return ROWS_FRAC;
            if (v < 0) return ROWS_NEG;
            if (v > INT32_MAX || (sum += v) > INT32_MAX) return ROWS_BIG;
            ++p;
        }
        prev = c2;
        if (sum > 0) f(c2 - x, sum);
    }
    return ROWS_OK;
}

inline int64_t align16(int64_t v) { return (v + 15) / 16 * 16; }

template <class F>
void rows_parallel(int64_t n, int n_threads, F&& body) {
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(nt, 32), n / 256 + 1));
    std::atomic<int64_t> next(0);
    const int64_t step = 128;
    auto work = [&]() {
        for (;;) {
            const int64_t x0 = next.fetch_add(step);
            if (x0 >= n) return;
            body(x0, std::min(n, x0 + step));
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int pk_rows_pack(const int64_t* bin1_offset, const void* bin2, int32_t bin2_bytes, int64_t bin2_base,
                            const void* count, int32_t count_kind, int64_t n_bins, int32_t nd_enc, int32_t far_mode,
                            const double* weights, void* out, int64_t capacity, int64_t* needed, int32_t n_threads) {
    if (!bin1_offset || n_bins < 0 || nd_enc < 1 || (bin2_bytes != 4 && bin2_bytes != 8) || count_kind < 0 || count_kind > 2 ||
        far_mode < 0 || far_mode > 1 || !needed || bin1_offset[0] < 0 || (bin1_offset[n_bins] > bin1_offset[0] && (!bin2 || !count))) {
        pk_set_error("pk_rows_pack: bad argument");
        return PK_EINVAL;
    }
    for (int64_t x = 0; x < n_bins; ++x)
        if (bin1_offset[x + 1] < bin1_offset[x]) {
            pk_set_error("pk_rows_pack: bin1_offset is not a row pointer (row %lld)", (long long)x);
            return PK_EINVAL;
        }
    const RowsIn in{bin1_offset, bin2, bin2_bytes, bin2_base, count, count_kind, n_bins, nd_enc};
    const int64_t n = n_bins, W = (nd_enc + 31) / 32;
    std::vector<RowTally> tally((size_t)n);
    std::atomic<int> bad(0);
    std::atomic<int64_t> bad_row(-1);
    // far_mode 1: a far pixel only matters as the witness that makes a bin `valid` (utils.py:146-156: a bin with any
    // finite pixel, k_band_rows' far loop); bins that a pixel of the bitmap section already makes valid need none.
    // The judgement is the device's: count > 0 and, with weights, isfinite((w_x w_y) count) (this file is compiled
    // without contraction: the same two roundings as __dmul_rn).
    std::vector<std::atomic<uint8_t>> wit(far_mode ? (size_t)n : 0);
    for (auto& a : wit) a.store(0, std::memory_order_relaxed);
    auto finite_px = [&](int64_t x, int64_t d, int64_t c) -> bool {
        return !weights || std::isfinite((weights[x] * weights[x + d]) * (double)c);
    };
    auto keep_far = [&](int64_t x, int64_t d, int64_t c) -> bool {
        if (!far_mode) return true;
        return finite_px(x, d, c) && !(wit[(size_t)x].load(std::memory_order_relaxed) && wit[(size_t)(x + d)].load(std::memory_order_relaxed));
    };
    rows_parallel(n, n_threads, [&](int64_t x0, int64_t x1) {
        for (int64_t x = x0; x < x1 && !bad.load(std::memory_order_relaxed); ++x) {
            RowTally t{0, 0, 0};
            const int rc = rows_walk(in, x, [&](int64_t d, int64_t c) {
                if (d < in.nd) {
                    ++t.band; t.esc += c >= 255;
                    if (far_mode && finite_px(x, d, c)) {
                        wit[(size_t)x].store(1, std::memory_order_relaxed);
                        wit[(size_t)(x + d)].store(1, std::memory_order_relaxed);
                    }
                } else ++t.far;
            });
            if (rc != ROWS_OK) { bad.store(rc); bad_row.store(x); return; }
            tally[(size_t)x] = t;
        }
    });
    if (bad.load()) {
        static const char* why[] = {"", "pixels below the diagonal (storage-mode is not symmetric-upper)", "pixels are not in cooler order (bin1, then bin2)",
                                    "negative pixel counts", "pixel counts outside int32", "non-integer pixel counts; the Poisson filter (scoreUtils.py:59-60) needs raw counts"};
        pk_set_error("pk_rows_pack: row %lld: %s", (long long)bad_row.load(), why[bad.load()]);
        return PK_EINVAL;
    }
    if (far_mode)                                            // the witnesses are complete: which far pixels stay
        rows_parallel(n, n_threads, [&](int64_t x0, int64_t x1) {
            for (int64_t x = x0; x < x1; ++x) {
                if (tally[(size_t)x].far == 0) continue;
                int64_t kept = 0;
                rows_walk(in, x, [&](int64_t d, int64_t c) { if (d >= in.nd && keep_far(x, d, c)) ++kept; });
                tally[(size_t)x].far = kept;
            }
        });
    // placement
    std::vector<int64_t> band_at((size_t)n + 1), esc_at((size_t)n + 1), far_at((size_t)n + 1);
    int64_t nb = 0, ne = 0, nf = 0;
    for (int64_t x = 0; x < n; ++x) {
        band_at[(size_t)x] = nb; esc_at[(size_t)x] = ne; far_at[(size_t)x] = nf;
        nb += tally[(size_t)x].band; ne += tally[(size_t)x].esc; nf += tally[(size_t)x].far;
    }
    band_at[(size_t)n] = nb; esc_at[(size_t)n] = ne; far_at[(size_t)n] = nf;
    if (nb >= (1LL << 32)) {
        pk_set_error("pk_rows_pack: more than 2^32 band pixels in one chromosome");
        return PK_EUNSUPPORTED;
    }
    const int64_t sizes[7] = {n * W * 4, (n + 1) * 4, nb, 3 * ne * 4, (n + 1) * 8, nf * 4, nf * 4};
    int64_t offs[7], at = 16 * 8;
    for (int i = 0; i < 7; ++i) { at = align16(at); offs[i] = at; at += sizes[i]; }
    const int64_t total = align16(at);
    *needed = total;
    if (!out || capacity < total) {
        if (out) pk_set_error("pk_rows_pack: the blob needs %lld bytes, the buffer holds %lld", (long long)total, (long long)capacity);
        return out ? PK_ECAPACITY : PK_OK;
    }
    uint8_t* blob = static_cast<uint8_t*>(out);
    const int64_t head[16] = {PK_ROWS_MAGIC, n, nd_enc, W, nb, ne, nf, offs[0], offs[1], offs[2], offs[3], offs[4], offs[5], offs[6], total, 0};
    memcpy(blob, head, sizeof head);
    // alignment gaps and the tail are zero, like the Python packer's
    int64_t end = 128;
    for (int i = 0; i < 7; ++i) { memset(blob + end, 0, (size_t)(offs[i] - end)); end = offs[i] + sizes[i]; }
    memset(blob + end, 0, (size_t)(total - end));
    uint32_t* bits = reinterpret_cast<uint32_t*>(blob + offs[0]);
    uint32_t* cnt_off = reinterpret_cast<uint32_t*>(blob + offs[1]);
    uint8_t* cnt8 = blob + offs[2];
    int32_t* esc = reinterpret_cast<int32_t*>(blob + offs[3]);
    int64_t* far_off = reinterpret_cast<int64_t*>(blob + offs[4]);
    int32_t* far_b2 = reinterpret_cast<int32_t*>(blob + offs[5]);
    int32_t* far_cnt = reinterpret_cast<int32_t*>(blob + offs[6]);
    cnt_off[n] = (uint32_t)nb;
    far_off[n] = nf;
    rows_parallel(n, n_threads, [&](int64_t x0, int64_t x1) {
        for (int64_t x = x0; x < x1; ++x) {
            uint32_t* brow = bits + x * W;
            for (int64_t k = 0; k < W; ++k) brow[k] = 0;
            int64_t ib = band_at[(size_t)x], ie = esc_at[(size_t)x], jf = far_at[(size_t)x];
            cnt_off[x] = (uint32_t)ib;
            far_off[x] = jf;
            rows_walk(in, x, [&](int64_t d, int64_t c) {
                if (d < in.nd) {
                    brow[d >> 5] |= 1u << (d & 31);
                    cnt8[ib++] = (uint8_t)std::min<int64_t>(c, 255);
                    if (c >= 255) { esc[ie] = (int32_t)x; esc[ne + ie] = (int32_t)d; esc[2 * ne + ie] = (int32_t)c; ++ie; }
                } else if (keep_far(x, d, c)) {
                    far_b2[jf] = (int32_t)(x + d);
                    far_cnt[jf] = (int32_t)c;
                    ++jf;
                }
            });
        }
    });
    return PK_OK;
}
