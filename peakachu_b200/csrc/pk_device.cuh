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

// packed forest node (see pk_common.cuh): leaf <=> sign bit of .y clear
#define PK_NODE_INTERNAL(y) ((int)(y) < 0)
#define PK_NODE_FEAT4(y) ((y) & 0xFFCu)               /* feature index * 4 (byte offset into a float row) */
#define PK_NODE_FEAT(y) (((y) & 0xFFCu) >> 2)
#define PK_NODE_MGL(y) (((y) >> 30) & 1u)
#define PK_NODE_ROFF(y) (((y) >> 12) & 0x3FFFFu)
#define PK_NODE_ROFF8(y) (((y) >> 9) & 0x1FFFF8u)     /* right-child offset * 8 (byte offset) */
