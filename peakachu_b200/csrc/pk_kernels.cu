// peakachu_b200: sm_100a kernels of the loop-scoring path.
//
// Data layout in HBM (per chromosome handle):
//   band[d][x]  int32 raw count of pixel (x, x+d), d in [0, upper+2w], row pitch a
//               multiple of 32 elements. Diagonal-major so that (1) the Poisson scan
//               over one distance and (2) a window row segment are both contiguous
//               in x, and per-diagonal sums stream linearly.
//   w[x]        float64 balancing weight; the balanced value of a pixel is computed on
//               the fly as (w[r] * w[c]) * count -- the same IEEE products as the
//               cooler boundary (coolio.py) -- never stored.
// All float64 arithmetic that feeds a feature or the expected curve uses explicit
// round-to-nearest intrinsics (__dadd_rn/__dmul_rn/__ddiv_rn) so that no FMA is
// contracted: results are bit-identical to numpy/scipy/numba on x86-64.
#include <math_constants.h>

#include <algorithm>

#include "pk_common.cuh"
#include "pk_device.cuh"

// ---------------------------------------------------------------------------
// K1  band build: scatter upper-triangle pixels into the diagonal-major band and
//     mark bins that own at least one finite pixel (utils.py:146-156).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter_pixels(
    const int32_t* __restrict__ b1, const int32_t* __restrict__ b2, const int32_t* __restrict__ cnt,
    long long nnz, const double* __restrict__ w, int n, int ND, long long pitch, int balanced,
    int32_t* __restrict__ band, uint8_t* __restrict__ valid, int32_t* __restrict__ flags) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    int x = b1[i], y = b2[i], c = cnt[i];
    if (x > y) { int t = x; x = y; y = t; }
    if (c == 0 || x < 0 || y >= n) return;
    bool fin = true;
    if (balanced) {
        double v = __dmul_rn(__dmul_rn(w[x], w[y]), (double)c);
        fin = isfinite(v);
    } else {
        fin = c > 0;     // raw mode: column sum > 0 (utils.py:148-149), counts are positive
    }
    if (fin) { valid[x] = 1; valid[y] = 1; }
    int d = y - x;
    if (d < ND) {
        atomicAdd(&band[(long long)d * pitch + x], c);
        if (c > flags[1]) atomicMax(&flags[1], c);
    }
}

// ---------------------------------------------------------------------------
// K4  window features (scoreUtils.py:70-93 + utils.py:180-237 + scipy gaussian_filter
//     + utils.image_normalize). One warp per candidate; the (2w+1)^2 window lives in
//     shared memory as float64.
//       gather  W[a][b] = M[x-w+a, y-w+b]  (symmetric; |col-row| >= upper+2w trimmed to 0)
//       reject  count_nonzero < 0.1*size | ll_mean <= 0 | centre/ll_mean <= 0.1
//               (ll_mean = sequential row-major sum of W[:w,:w] / w^2, numba order)
//       normalise W / exp[|col-row|]
//       gaussian sigma=1: axis 0 then axis 1, radius 4, reflect,
//               t = x[c]*k4; t += (x[c-4]+x[c+4])*k0; ... ; t += (x[c-1]+x[c+1])*k3
//       min-max (G - min) / (max - min), NaN propagates; cast to float32
// ---------------------------------------------------------------------------
#define PK_FEAT_WARPS 4

__global__ void __launch_bounds__(PK_FEAT_WARPS * 32) k_features(
    const int32_t* __restrict__ band, const double* __restrict__ w, const double* __restrict__ expv,
    int n, long long pitch, int balanced, int wd, int ND,
    const int32_t* __restrict__ cx, const int32_t* __restrict__ cd, const int32_t* __restrict__ crank,
    long long n_cand, uint8_t* __restrict__ keep, float* __restrict__ fea32, double* __restrict__ fea64,
    int32_t* __restrict__ batch_win, unsigned long long* __restrict__ counters) {
    extern __shared__ double smem[];
    const int S = 2 * wd + 1, F = S * S;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* A = smem + (size_t)wib * 2 * F;
    double* T = A + F;
    const long long warp0 = (long long)blockIdx.x * PK_FEAT_WARPS + wib;
    const long long nwarps = (long long)gridDim.x * PK_FEAT_WARPS;
    for (long long i = warp0; i < n_cand; i += nwarps) {
        const int x = cx[i], d = cd[i], y = x + d;
        bool ok = (x - wd >= 0) && (y + wd + 1 <= n);        // scoreUtils.py:75
        int nz = 0;
        if (ok) {
            for (int idx = lane; idx < F; idx += 32) {
                int a = idx / S, b = idx - a * S;
                int r = x - wd + a, c = y - wd + b;
                int dd = c - r, ad = dd < 0 ? -dd : dd;
                int lo = dd < 0 ? c : r;
                double v = 0.0;
                if (ad < ND - 1) {                           // scoreUtils.py:31: col-row < upper+2w
                    int cnt = band[(long long)ad * pitch + lo];
                    v = pk_value(cnt, balanced ? w[r] : 0.0, balanced ? w[c] : 0.0, balanced);
                }
                A[idx] = v;
                nz += (v != 0.0);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
            __syncwarp();
            if ((double)nz < (double)F * 0.1) ok = false;    // utils.py:225
        }
        if (ok) {
            double s = 0.0;                                  // utils.py:228, numba sequential mean
            for (int a = 0; a < wd; ++a)
                for (int b = 0; b < wd; ++b) s = __dadd_rn(s, A[a * S + b]);
            double ll = __ddiv_rn(s, (double)(wd * wd));
            ok = (ll > 0.0) && (__ddiv_rn(A[wd * S + wd], ll) > 0.1);   // utils.py:229-232
        }
        if (!ok) {
            if (lane == 0) keep[i] = 0;
            __syncwarp();
            continue;
        }
        // distance normalisation (utils.py:187-200)
        for (int idx = lane; idx < F; idx += 32) {
            int a = idx / S, b = idx - a * S;
            int dd = (y - wd + b) - (x - wd + a);
            int ad = dd < 0 ? -dd : dd;
            A[idx] = __ddiv_rn(A[idx], expv[ad]);
        }
        __syncwarp();
        // gaussian, axis 0 (rows a)
        for (int idx = lane; idx < F; idx += 32) {
            int a = idx / S, b = idx - a * S;
            double t = __dmul_rn(A[idx], PK_GK[4]);
#pragma unroll
            for (int j = 4; j >= 1; --j) {
                double p = A[pk_reflect(a - j, S) * S + b], q = A[pk_reflect(a + j, S) * S + b];
                t = __dadd_rn(t, __dmul_rn(__dadd_rn(p, q), PK_GK[4 - j]));
            }
            T[idx] = t;
        }
        __syncwarp();
        // gaussian, axis 1 (columns b) + min/max
        double mn = CUDART_INF, mx = -CUDART_INF;
        bool has_nan = false;
        for (int idx = lane; idx < F; idx += 32) {
            int a = idx / S, b = idx - a * S;
            double t = __dmul_rn(T[idx], PK_GK[4]);
#pragma unroll
            for (int j = 4; j >= 1; --j) {
                double p = T[a * S + pk_reflect(b - j, S)], q = T[a * S + pk_reflect(b + j, S)];
                t = __dadd_rn(t, __dmul_rn(__dadd_rn(p, q), PK_GK[4 - j]));
            }
            A[idx] = t;          // A is free again (all lanes passed the previous __syncwarp)
            has_nan |= isnan(t);
            mn = fmin(mn, t); mx = fmax(mx, t);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        has_nan = __any_sync(0xffffffffu, has_nan);
        if (has_nan) { mn = CUDART_NAN; mx = CUDART_NAN; }   // numba min/max return NaN
        const double range = __dsub_rn(mx, mn);
        __syncwarp();
        for (int idx = lane; idx < F; idx += 32) {
            double v = __ddiv_rn(__dsub_rn(A[idx], mn), range);   // utils.py:207
            if (fea64) fea64[i * F + idx] = v;
            fea32[i * F + idx] = __double2float_rn(v);
        }
        if (lane == 0) {
            keep[i] = 1;
            atomicAdd(&batch_win[crank[i] / PK_BATCH], 1);
            atomicAdd(&counters[1], 1ULL);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// K5  forest traversal (sklearn _apply_dense + predict_proba). One row per thread.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double pk_tree_eval(const uint2* __restrict__ nodes, uint32_t root,
                                               const float* __restrict__ xrow, uint32_t* leaf_slot) {
    uint32_t p = root;
    uint2 nd = nodes[p];
    while (PK_NODE_INTERNAL(nd.y)) {
        const float xv = xrow[PK_NODE_FEAT(nd.y)];
        const bool left = isnan(xv) ? (PK_NODE_MGL(nd.y) != 0u) : (xv <= __uint_as_float(nd.x));
        p = left ? p + 1u : p + PK_NODE_ROFF(nd.y);
        nd = nodes[p];
    }
    *leaf_slot = p;
    return __hiloint2double((int)nd.y, (int)nd.x);
}

__global__ void __launch_bounds__(128) k_forest(
    const uint2* __restrict__ nodes, const uint32_t* __restrict__ roots, const int32_t* __restrict__ orig,
    int n_trees, int F, const float* __restrict__ X, const uint8_t* __restrict__ keep, long long n_rows,
    int32_t* __restrict__ leaves, double* __restrict__ proba) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    if (keep && !keep[i]) return;
    const float* xrow = X + i * F;
    double acc = 0.0;
    for (int t = 0; t < n_trees; ++t) {
        uint32_t slot;
        double v = pk_tree_eval(nodes, roots[t], xrow, &slot);
        acc = __dadd_rn(acc, v);
        if (leaves) leaves[i * n_trees + t] = orig[slot];
    }
    if (proba) proba[i] = __ddiv_rn(acc, (double)n_trees);
}

// ---------------------------------------------------------------------------
// K6  emit: keep prob > min_prob (strict, scoreUtils.py:110) and, for whole
//     chromosomes, drop batches with <= 1 surviving window (scoreUtils.py:108).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_emit(
    const int32_t* __restrict__ cx, const int32_t* __restrict__ cd, const int32_t* __restrict__ crank,
    const uint8_t* __restrict__ keep, const double* __restrict__ prob, const long long* __restrict__ ncand_dev,
    long long cap, double thre, const int32_t* __restrict__ batch_win, int apply_rule,
    const int32_t* __restrict__ band, const double* __restrict__ w, long long pitch, int balanced,
    int32_t* __restrict__ rx, int32_t* __restrict__ ry, double* __restrict__ rp, double* __restrict__ rv,
    int32_t* __restrict__ rb, int32_t* __restrict__ rowcnt, int32_t* __restrict__ rrank,
    unsigned long long* __restrict__ counters) {
    const long long n_cand = min(ncand_dev[0], cap);
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_round = (n_cand + 31) / 32 * 32;        // whole warps take part in the ballot
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool out = false;
        int x = 0, d = 0, b = 0;
        double p = 0.0;
        if (i < n_cand) {
            // independent loads first: one memory round trip for the five columns
            const uint8_t kp = keep[i];
            const double pv = prob[i];
            const int rk = crank[i];
            x = cx[i]; d = cd[i];
            if (kp) {
                p = pv;
                b = rk / PK_BATCH;
                out = (p > thre) && (!apply_rule || batch_win[b] > 1);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, out);
        if (bal == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&counters[0], (unsigned long long)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (out) {
            const unsigned long long o = base + __popc(bal & ((1u << lane) - 1u));
            rx[o] = x; ry[o] = x + d; rp[o] = p; rb[o] = b;
            rrank[o] = atomicAdd(&rowcnt[x], 1);              // arrival number inside row x (pk_sort.cu)
            rv[o] = pk_value(band[(long long)d * pitch + x], balanced ? w[x] : 0.0, balanced ? w[x + d] : 0.0, balanced);
        }
    }
}

// ---------------------------------------------------------------------------
// host-side launchers (called from pk_api.cpp)
// ---------------------------------------------------------------------------
int pk_launch_scatter(pk_chrom* c, const int32_t* b1, const int32_t* b2, const int32_t* cnt, int64_t nnz) {
    if (nnz == 0) return PK_OK;
    unsigned grid = (unsigned)((nnz + 255) / 256);
    k_scatter_pixels<<<grid, 256, 0, c->stream>>>(b1, b2, cnt, nnz, c->d_w, c->n, c->ND, c->pitch,
                                                 c->balanced, c->d_band, c->d_valid, c->d_flags);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int pk_launch_features(pk_chrom* c, double* d_fea64) {
    if (c->n_cand == 0) return PK_OK;
    size_t smem = (size_t)PK_FEAT_WARPS * 2 * c->F * sizeof(double);
    PK_OPT_IN_SMEM(k_features, smem, c->device);
    long long want = (c->n_cand + PK_FEAT_WARPS - 1) / PK_FEAT_WARPS;
    unsigned grid = (unsigned)std::min<long long>(want, 148LL * 16);
    k_features<<<grid, PK_FEAT_WARPS * 32, smem, c->stream>>>(c->d_band, c->d_w, c->d_exp, c->n, c->pitch, c->balanced, c->w,
                                                            c->ND, c->d_cx, c->d_cd, c->d_crank, c->n_cand, c->d_keep,
                                                            c->d_fea32, d_fea64, c->d_batch_win, c->d_counters);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int pk_launch_forest(const pk_forest* f, const float* X, const uint8_t* keep, int64_t n_rows, int32_t* leaves,
                     double* proba, cudaStream_t stream) {
    if (n_rows == 0) return PK_OK;
    unsigned grid = (unsigned)((n_rows + 127) / 128);
    k_forest<<<grid, 128, 0, stream>>>(f->d_nodes, f->d_root, f->d_orig, f->n_trees, f->n_features, X, keep, n_rows,
                                       leaves, proba);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int pk_launch_emit(pk_chrom* c, double thre) {
    k_emit<<<148 * 4, 256, 0, c->stream>>>(c->d_cx, c->d_cd, c->d_crank, c->d_keep, c->d_prob, c->d_ncand, c->cand_cap, thre,
                                           c->d_batch_win, c->whole ? 1 : 0, c->d_band, c->d_w, c->pitch, c->balanced,
                                           c->d_rx, c->d_ry, c->d_rp, c->d_rv, c->d_rb, c->d_rowcnt, c->d_rrank, c->d_counters);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}
