// Device helpers shared by the kernels of peakachu_b200.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

// balanced value (w[r] * w[c]) * count; non-finite pixels are trimmed (scoreUtils.py:31)
__device__ __forceinline__ double pk_value(int cnt, double wr, double wc, int balanced) {
    if (cnt == 0) return 0.0;
    if (!balanced) return (double)cnt;
    double v = __dmul_rn(__dmul_rn(wr, wc), (double)cnt);
    return isfinite(v) ? v : 0.0;
}

// scipy.ndimage gaussian_filter(sigma=1): exp(-x^2/2)/sum, x=-4..4; [0]=far tap .. [4]=centre
__device__ constexpr double PK_GK[5] = {0x1.18a9c4fd536c6p-13, 0x1.22724cb7eb269p-8, 0x1.ba4b99d1799abp-5,
                                        0x1.ef8eb9ad499bap-3, 0x1.9884a307594fbp-2};

// scipy mode='reflect' (d c b a | a b c d | d c b a)
__host__ __device__ constexpr int pk_reflect(int i, int S) { return i < 0 ? -i - 1 : (i >= S ? 2 * S - i - 1 : i); }

// a / b, correctly rounded, from a correctly rounded reciprocal r = RN(1/b): two
// residual corrections (q' = q + (a - b q) r with exact FMA residuals). The first makes q
// faithful, the second is then correctly rounded (Markstein). Valid away from
// overflow/underflow; callers take __ddiv_rn otherwise. Checked against __ddiv_rn on the
// device by pk_selftest_divide.
__device__ __forceinline__ double pk_div_r(double a, double b, double r) {
    double q = __dmul_rn(a, r);
    q = __fma_rn(__fma_rn(-b, q, a), r, q);
    q = __fma_rn(__fma_rn(-b, q, a), r, q);
    return q;
}
__device__ __forceinline__ bool pk_div_safe(double v) { return v >= 1e-100 && v <= 1e100; }   // false for NaN

// packed forest node (see pk_common.cuh): leaf <=> sign bit of .y clear
#define PK_NODE_INTERNAL(y) ((int)(y) < 0)
#define PK_NODE_FEAT4(y) ((y) & 0xFFCu)               /* feature index * 4 (byte offset into a float row) */
#define PK_NODE_FEAT(y) (((y) & 0xFFCu) >> 2)
#define PK_NODE_MGL(y) (((y) >> 30) & 1u)
#define PK_NODE_ROFF(y) (((y) >> 12) & 0x3FFFFu)
#define PK_NODE_ROFF8(y) (((y) >> 9) & 0x1FFFF8u)     /* right-child offset * 8 (byte offset) */

// Exclusive scan of in[0..m) into out[0..m) by ONE CTA of 1024 threads, out[m] = total (returned to
// every thread). `in` and `out` are 16-byte aligned 32-bit arrays. Tiles of 4096 values: a thread
// moves its four consecutive values with one 128-bit load / store (coalesced -- a single SM's
// load/store unit is the bottleneck of a one-CTA kernel), the next tile's load is issued before
// the current tile is scanned, and a shuffle scan of the per-thread sums supplies the offsets.
// `s_warp` is 33 words of shared memory.
template <typename T>
__device__ __forceinline__ uint32_t pk_cta_scan_1024(const T* __restrict__ in, long long m, T* __restrict__ out,
                                                    uint32_t* s_warp) {
    static_assert(sizeof(T) == 4, "32-bit elements");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    auto load4 = [&](long long i0) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (i0 + 3 < m) v = *reinterpret_cast<const uint4*>(in + i0);
        else {
            if (i0 < m) v.x = (uint32_t)in[i0];
            if (i0 + 1 < m) v.y = (uint32_t)in[i0 + 1];
            if (i0 + 2 < m) v.z = (uint32_t)in[i0 + 2];
        }
        return v;
    };
    uint32_t carry = 0;
    uint4 nxt = load4((long long)tid * 4);
    for (long long base = 0; base < m; base += 4096) {
        const long long i0 = base + (long long)tid * 4;
        const uint4 v = nxt;
        if (base + 4096 < m) nxt = load4(i0 + 4096);
        const uint32_t sum = v.x + v.y + v.z + v.w;
        uint32_t x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane], y = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= o) y += t;
            }
            s_warp[lane] = y - w;                 // exclusive warp offsets
            if (lane == 31) s_warp[32] = y;       // tile total
        }
        __syncthreads();
        uint4 r;
        r.x = carry + s_warp[wid] + x - sum;
        r.y = r.x + v.x; r.z = r.y + v.y; r.w = r.z + v.z;
        if (i0 + 3 < m) *reinterpret_cast<uint4*>(out + i0) = r;
        else {
            if (i0 < m) out[i0] = (T)r.x;
            if (i0 + 1 < m) out[i0 + 1] = (T)r.y;
            if (i0 + 2 < m) out[i0 + 2] = (T)r.z;
        }
        carry += s_warp[32];
        __syncthreads();                          // s_warp is reused by the next tile
    }
    if (tid == 0) out[m] = (T)carry;
    return carry;
}
