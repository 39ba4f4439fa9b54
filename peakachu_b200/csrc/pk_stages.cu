// peakachu_b200: the stages in front of the scoring kernel -- band build from CSR
// pixels, per-diagonal sums, expected-curve fit, Poisson candidate scan. All of them
// run back to back on the handle's stream; none needs the host.
#include "pk_common.cuh"
#include "pk_device.cuh"

#include <algorithm>

// ---------------------------------------------------------------------------
// S1  band build from CSR-ordered pixels (cooler order: sorted by bin1, then bin2).
//     A CTA owns 32 consecutive rows x0..x0+31. Each warp streams whole rows (its
//     pixels are contiguous: coalesced loads), scatters counts into a shared-memory
//     tile T[d][x - x0] and the CTA then writes the tile out as full 128-byte lines,
//     zeros included -- the band needs no memset and no global atomics.
//     Side products: valid[] (utils.py:146-156) and the largest in-band count.
// ---------------------------------------------------------------------------
// TY / TC: element types of the column and count arrays. DELTA: the column array holds
// bin2 - bin1 (pk_chrom_upload_csr16) instead of bin2.
// Write a CTA's tile T[d][0..R) (R <= 64 rows, pitch TP) to the diagonal-major band: a warp per distance,
// lanes along the rows, pointers advanced by addition (no 64-bit multiply per store).
__device__ __forceinline__ void pk_tile_writeout(const int32_t* __restrict__ s_tile, int TP, int R, int32_t* __restrict__ band,
                                                 long long pitch, int x0, int n, int ND, int lane, int wib) {
    const bool a = lane < R && x0 + lane < n, b = lane + 32 < R && x0 + lane + 32 < n;
    int32_t* dst = band + (long long)wib * pitch + x0 + lane;
    const int32_t* src = s_tile + wib * TP + lane;
    const long long dstep = 8 * pitch;
    const int sstep = 8 * TP;
    for (int d = wib; d < ND; d += 8) {
        if (a) dst[0] = src[0];
        if (b) dst[32] = src[32];
        dst += dstep; src += sstep;
    }
}

// exponent of a weight inside [2^-150, 2^150] (false for 0, denormals, inf, NaN): products of two such
// weights and a 31-bit count are finite, so the finiteness test of a balanced pixel needs no arithmetic
__device__ __forceinline__ bool pk_weight_tame(double w) {
    const unsigned e = ((unsigned)__double2hiint(w) >> 20) & 0x7FFu;
    return e >= 1023u - 150u && e <= 1023u + 150u;
}

template <typename TY, typename TC, bool DELTA>
__global__ void __launch_bounds__(256) k_band_csr(
    const long long* __restrict__ rowptr, const TY* __restrict__ b2, const TC* __restrict__ cnt,
    const double* __restrict__ w, int n, int ND, long long pitch, int balanced, int R,
    int32_t* __restrict__ band, uint8_t* __restrict__ valid, int32_t* __restrict__ flags) {
    extern __shared__ int32_t s_tile[];                 // [ND][R + 1], R rows per CTA
    const int TP = R + 1;
    const int x0 = blockIdx.x * R;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    for (int i = tid; i < ND * TP; i += 256) s_tile[i] = 0;
    __syncthreads();
    int cmax = 0;
    for (int xl = wib; xl < R; xl += 8) {
        const int x = x0 + xl;
        if (x >= n) break;
        const long long p0 = rowptr[x], p1 = rowptr[x + 1];
        const double wx = balanced ? w[x] : 0.0;
        const bool tame_x = !balanced || pk_weight_tame(wx);
        bool any = false;
        for (long long pb = p0; pb < p1; pb += 128) {
            int y[4], c[4];
            double wy[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long p = pb + j * 32 + lane;
                y[j] = -1; c[j] = 0;
                if (p < p1) { y[j] = DELTA ? x + (int)b2[p] : (int)b2[p]; c[j] = (int)cnt[p]; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c[j] == 0 || y[j] < x || y[j] >= n) c[j] = 0;
                wy[j] = (balanced && c[j] != 0) ? w[y[j]] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c[j] == 0) continue;
                bool fin;
                if (!balanced) fin = c[j] > 0;
                else if (tame_x && pk_weight_tame(wy[j])) fin = true;
                else fin = isfinite(__dmul_rn(__dmul_rn(wx, wy[j]), (double)c[j]));
                if (fin) { any = true; valid[y[j]] = 1; }
                const int d = y[j] - x;
                if (d < ND) {
                    atomicAdd(&s_tile[d * TP + xl], c[j]);     // duplicates are summed like utils.tocsr
                    cmax = max(cmax, c[j]);
                }
            }
        }
        if (__any_sync(0xffffffffu, any) && lane == 0) valid[x] = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    if (lane == 0 && cmax > 0) atomicMax(&flags[1], cmax);
    __syncthreads();
    pk_tile_writeout(s_tile, TP, R, band, pitch, x0, n, ND, lane, wib);
}

// ---------------------------------------------------------------------------
// S1'  band build from packed pixel rows (pk_chrom_upload_rows; layout in include/peakachu_b200.h and
//      peakachu_b200/rowpack.py): per row a presence bitmap over the first nd_enc distances, one count
//      byte per present pixel (255 = escaped to a side list), far pixels (d >= nd_enc) as CSR columns.
//      Same tiling as k_band_csr: a CTA owns R rows, a warp walks a row -- lane l decodes bitmap word l,
//      i.e. distances 32 l .. 32 l + 31 -- and the CTA writes the tile as full lines. The bitmap makes
//      duplicates impossible, so the tile is filled with plain stores. Far pixels only mark `valid`.
// ---------------------------------------------------------------------------
struct PkRowsView {
    const uint32_t* bits; const uint32_t* cnt_off; const uint8_t* cnt8;
    const int32_t* esc;            // [3][n_esc]: x | d | count
    long long n_esc;
    const long long* far_off; const int32_t* far_b2; const int32_t* far_cnt;
    int nd_enc, W;
};

__global__ void __launch_bounds__(256) k_band_rows(
    const PkRowsView v, const double* __restrict__ w, int n, int ND, long long pitch, int balanced, int R,
    int32_t* __restrict__ band, uint8_t* __restrict__ valid, int32_t* __restrict__ flags) {
    extern __shared__ int32_t s_tile[];                 // [ND][R + 1]
    const int TP = R + 1;
    const int x0 = blockIdx.x * R;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    for (int i = tid; i < ND * TP; i += 256) s_tile[i] = 0;
    __syncthreads();
    int cmax = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int xl = wib; xl < R; xl += 8) {
        const int x = x0 + xl;
        if (x >= n) break;
        const double wx = balanced ? w[x] : 0.0;
        const bool tame_x = !balanced || pk_weight_tame(wx);
        bool any = false;
        const uint8_t* cb = v.cnt8 + v.cnt_off[x];
        int carry = 0;
        for (int w0 = 0; w0 < v.W; w0 += 32) {
            // lane l holds bitmap word w0 + l and the number of pixels in the words before it
            const uint32_t mine = (w0 + lane < v.W) ? v.bits[(size_t)x * v.W + w0 + lane] : 0u;
            const int pc = __popc(mine);
            int pre = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
            pre += carry - pc;                                             // exclusive, from the start of the row
            carry = __shfl_sync(0xffffffffu, pre + pc, 31);
            const int nw = min(32, v.W - w0);
            // one word per step: lane = distance inside the word, so count bytes and weights are read
            // coalesced; eight words per round, every load of a round issued before the first use
            for (int q0 = 0; q0 < nw; q0 += 8) {
                int c[8];
                double wy[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = min(q0 + j, 31);
                    const uint32_t b = __shfl_sync(0xffffffffu, mine, q);
                    const int base = __shfl_sync(0xffffffffu, pre, q);
                    const int y = x + (w0 + q) * 32 + lane;
                    const bool have = q0 + j < nw && ((b >> lane) & 1u) && y < n;
                    c[j] = have ? (int)cb[base + __popc(b & lt)] : 0;
                    wy[j] = (have && balanced) ? w[y] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (c[j] == 0 || c[j] == 255) continue;                 // 255: an escaped count (k_band_escapes)
                    const int d = (w0 + q0 + j) * 32 + lane;
                    bool fin = true;
                    if (balanced && !(tame_x && pk_weight_tame(wy[j]))) fin = isfinite(__dmul_rn(__dmul_rn(wx, wy[j]), (double)c[j]));
                    if (fin) { any = true; valid[x + d] = 1; }
                    if (d < ND) { s_tile[d * TP + xl] = c[j]; cmax = max(cmax, c[j]); }
                }
            }
        }
        // far pixels: never in the band, but they make their bins valid (utils.py:146-156)
        for (long long p = v.far_off[x] + lane; p < v.far_off[x + 1]; p += 32) {
            const int y = v.far_b2[p], c = v.far_cnt[p];
            if (c <= 0 || y < x || y >= n) continue;
            bool fin = true;
            if (balanced) fin = isfinite(__dmul_rn(__dmul_rn(wx, w[y]), (double)c));
            if (fin) { any = true; valid[y] = 1; }
        }
        if (__any_sync(0xffffffffu, any) && lane == 0) valid[x] = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    if (lane == 0 && cmax > 0) atomicMax(&flags[1], cmax);
    __syncthreads();
    pk_tile_writeout(s_tile, TP, R, band, pitch, x0, n, ND, lane, wib);
}

// ---------------------------------------------------------------------------
// S1''  row-major copy of the band (pk_chrom.d_band2) for the fused kernel's TMA window fetch: a tiled
//       transpose band[o][r] -> band2[r][o], 32 x 32 cells per CTA step, coalesced on both sides (the band
//       was just written and sits in L2). Columns o >= ND - 1 (the trimmed diagonal, padding) are zeroed.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_band_rowmajor(const int32_t* __restrict__ band, long long pitch, int n, int ND,
                                                       int32_t* __restrict__ band2, int P2) {
    __shared__ int32_t t[32][33];
    const int r0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int o0 = 0; o0 < P2; o0 += 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int o = o0 + ty + 8 * k, r = r0 + tx;
            t[ty + 8 * k][tx] = (o < ND - 1 && r < n) ? band[(long long)o * pitch + r] : 0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + ty + 8 * k, o = o0 + tx;
            if (r < n && o < P2) band2[(long long)r * P2 + o] = t[tx][ty + 8 * k];
        }
        __syncthreads();
    }
}

int pk_launch_band_rowmajor(pk_chrom* c) {
    k_band_rowmajor<<<(unsigned)((c->n + 31) / 32), 256, 0, c->stream>>>(c->d_band, c->pitch, c->n, c->ND, c->d_band2, (int)c->P2);
    PK_CUDA(cudaGetLastError());
    c->band2_valid = true;
    return PK_OK;
}

// escaped counts (>= 255) of the packed rows: a few per row at most, written after the tiles
__global__ void __launch_bounds__(256) k_band_escapes(const PkRowsView v, const double* __restrict__ w, int n, int ND,
                                                      long long pitch, int balanced, int32_t* __restrict__ band,
                                                      uint8_t* __restrict__ valid, int32_t* __restrict__ flags) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= v.n_esc) return;
    const int x = v.esc[i], d = v.esc[v.n_esc + i], c = v.esc[2 * v.n_esc + i];
    if (x < 0 || d < 0 || x + d >= n || c <= 0) return;
    bool fin = true;
    if (balanced) fin = isfinite(__dmul_rn(__dmul_rn(w[x], w[x + d]), (double)c));
    if (fin) { valid[x] = 1; valid[x + d] = 1; }
    if (d < ND) {
        band[(long long)d * pitch + x] = c;
        if (c > flags[1]) atomicMax(&flags[1], c);
    }
}

// depth over packed rows: band part (a warp per row, lane per bitmap word), escapes, far pixels
__global__ void __launch_bounds__(256) k_depth_packed(const PkRowsView v, int n, int min_dis,
                                                      unsigned long long* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * 256) >> 5, warp = (blockIdx.x * 256 + threadIdx.x) >> 5;
    unsigned long long sum = 0;
    for (int x = warp; x < n; x += warps) {
        const uint8_t* cb = v.cnt8 + v.cnt_off[x];
        int carry = 0;
        for (int w0 = 0; w0 < v.W; w0 += 32) {
            const int wi = w0 + lane;
            uint32_t b = wi < v.W ? v.bits[(size_t)x * v.W + wi] : 0u;
            const int pc = __popc(b);
            int pre = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
            const uint8_t* mine = cb + carry + pre - pc;
            carry += __shfl_sync(0xffffffffu, pre, 31);
            for (int k = 0; b; ++k) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                const int c = mine[k], d = wi * 32 + bit;
                if (c != 255 && d >= min_dis && x + d < n) sum += (unsigned long long)c;
            }
        }
        for (long long p = v.far_off[x] + lane; p < v.far_off[x + 1]; p += 32) {
            const int y = v.far_b2[p], c = v.far_cnt[p];
            if (c > 0 && y - x >= min_dis && y < n) sum += (unsigned long long)c;
        }
    }
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < v.n_esc; i += (long long)gridDim.x * 256) {
        const int x = v.esc[i], d = v.esc[v.n_esc + i], c = v.esc[2 * v.n_esc + i];
        if (c > 0 && d >= min_dis && x >= 0 && x + d < n) sum += (unsigned long long)c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0 && sum) atomicAdd(total, sum);
}

// rowptr for pixels sorted by (bin1, bin2): rowptr[x] = first pixel with bin1 >= x
__global__ void __launch_bounds__(256) k_rowptr(const int32_t* __restrict__ b1, long long nnz, int n,
                                                long long* __restrict__ rowptr) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x > n) return;
    long long lo = 0, hi = nnz;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (b1[mid] < x) lo = mid + 1; else hi = mid;
    }
    rowptr[x] = lo;
}

// sortedness check of COO pixels (bin1 non-decreasing, bin1 <= bin2); flags[3] |= 1 if not
__global__ void __launch_bounds__(256) k_check_sorted(const int32_t* __restrict__ b1, const int32_t* __restrict__ b2,
                                                      long long nnz, int32_t* __restrict__ flags) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= nnz) return;
    bool bad = b1[i] > b2[i];
    if (i + 1 < nnz) bad |= b1[i] > b1[i + 1];
    if (bad) atomicOr(&flags[3], 1);
}

// valid[] bytes -> one bit per bin (word i covers bins 32i .. 32i+31); two padding words of 0
// Also raises flags[2] bit 2 when a finite non-zero weight lies outside [1e-45, 1e45]: balanced
// values then stay inside the range where the feature kernel's reciprocal division is exact.
__global__ void __launch_bounds__(256) k_valid_bits(const uint8_t* __restrict__ valid, const double* __restrict__ w,
                                                    int balanced, int n, uint32_t* __restrict__ vbits, int n_words,
                                                    int32_t* __restrict__ flags) {
    // one warp per word, lane = bit: coalesced reads of valid[] and w[]
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * 256) >> 5;
    bool wild = false;
    for (int i = (blockIdx.x * 256 + threadIdx.x) >> 5; i < n_words; i += warps) {
        const int x = i * 32 + lane;
        bool v = false;
        if (x < n) {
            v = valid[x] != 0;
            if (balanced) {
                const double a = fabs(w[x]);
                wild |= isfinite(a) && a != 0.0 && !(a >= 1e-45 && a <= 1e45);
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, v);
        if (lane == 0) vbits[i] = m;
    }
    if (wild) atomicOr(&flags[2], 4);
}

// ---------------------------------------------------------------------------
// S2  per-diagonal sums in numpy's pairwise order (utils.py:160-170).
//     One CTA per distance d. Each warp owns a contiguous slice of the diagonal:
//     it counts its valid pairs (valid[x] & valid[x+d]), a block scan turns the
//     counts into offsets, and the warp writes value(x, d) of its valid pairs,
//     zeros included and order kept, into the compacted scratch row. numpy's
//     pairwise tree (n > 128 -> n2 = n/2 - (n/2)%8 | rest; leaves of <= 128
//     elements with 8 strided accumulators combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
//     then a sequential tail) is evaluated leaf-parallel, combined by one thread.
// ---------------------------------------------------------------------------
#define PK_DS_THREADS 256
#define PK_DS_LEVELS 14

// numpy's leaf: eight strided accumulators r_j = a[j] + a[j+8] + ..., combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n%8 tail added one by one. Eight lanes
// (an aligned octet of a warp) take one accumulator each: all of a lane's values are
// loaded before the first add, so a leaf costs one memory latency, not sixteen.
// Every lane of the warp must call it; lanes of an octet share (a, n); n <= 128.
__device__ __forceinline__ double pk_leaf_sum8(const double* __restrict__ a, int n, int j /* lane & 7 */) {
    const int nb = n < 8 ? 0 : (n - (n % 8));      // elements covered by the accumulators
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (j + 8 * i < nb) ? a[j + 8 * i] : 0.0;
    double tail[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) tail[i] = (j == 0 && nb + i < n) ? a[nb + i] : 0.0;
    double r = v[0];
#pragma unroll
    for (int i = 1; i < 16; ++i)
        if (j + 8 * i < nb) r = __dadd_rn(r, v[i]);
    // pairwise combine across the octet (addition is commutative, so the partner order is free)
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    if (n < 8) r = 0.0;                              // numpy: plain loop from 0.0 for short arrays
#pragma unroll
    for (int i = 0; i < 7; ++i)
        if (nb + i < n) r = __dadd_rn(r, tail[i]);
    return r;                                         // valid in lane j == 0
}

// numpy's recursion, one level at a time, all threads: segments (start, size) of level l
// become those of level l+1 (a segment of more than 128 elements splits into
// n2 = m/2 - (m/2)%8 and the rest, others are carried over), order preserved. The split
// bitmask of every level is kept so that the sums can be combined back up the same tree.
template <int CAP>                 // leaves per diagonal handled in shared memory (n up to ~57 * CAP bins)
struct DiagSmem {
    int32_t seg_s[2][CAP];
    int32_t seg_m[2][CAP];
    double val[2][CAP];
    uint32_t split[PK_DS_LEVELS][CAP / 32];
    uint16_t wpre[PK_DS_LEVELS][CAP / 32];      // splits in the words before this one
    int32_t nseg[PK_DS_LEVELS + 1];
    int32_t wcnt[PK_DS_THREADS / 32];
};

// pair mask of bins x0..x0+31 on diagonal d: valid[x] & valid[x+d], from the bit vector
__device__ __forceinline__ uint32_t pk_pair_word(const uint32_t* __restrict__ vb, int x0, int d) {
    const int y0 = x0 + d;
    const uint32_t lo = vb[y0 >> 5], hi = vb[(y0 >> 5) + 1];
    return vb[x0 >> 5] & __funnelshift_r(lo, hi, y0 & 31);
}

// Compaction, wide: CTA (slice, d) owns PK_DC_SLICE consecutive x of diagonal d. Its output offset is the
// number of valid pairs before its slice, counted straight from the bit vector (at most n / 32 words), so
// no scan kernel and no ordering between CTAs is needed. Slice 0 also writes the diagonal's total.
#define PK_DC_SLICE 2048
__global__ void __launch_bounds__(256) k_diag_compact(
    const int32_t* __restrict__ band, const double* __restrict__ w, const uint32_t* __restrict__ vbits, int n_words,
    int n, long long pitch, int balanced, double* __restrict__ scratch, long long* __restrict__ out_cnt) {
    extern __shared__ uint32_t s_vbc[];                 // [n_words + 2]
    __shared__ int s_part[8];
    const int d = blockIdx.y, len = n - d;
    const int xs = blockIdx.x * PK_DC_SLICE;
    if (xs >= len && blockIdx.x != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < n_words + 2; i += 256) s_vbc[i] = i < n_words ? vbits[i] : 0u;
    __syncthreads();
    // valid pairs in the words before this slice (slice 0: in the whole diagonal, for out_cnt)
    const int w_end = blockIdx.x == 0 ? (len + 31) / 32 : xs / 32;
    int cntw = 0;
    for (int wi = tid; wi < w_end; wi += 256) {
        uint32_t m = pk_pair_word(s_vbc, wi * 32, d);
        if (wi * 32 + 32 > len) m &= (len - wi * 32 > 0) ? ((1u << (len - wi * 32)) - 1u) : 0u;
        cntw += __popc(m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cntw += __shfl_xor_sync(0xffffffffu, cntw, o);
    if (lane == 0) s_part[wid] = cntw;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) before += s_part[k];
    if (blockIdx.x == 0) {
        if (tid == 0) out_cnt[d] = before;
        before = 0;
    }
    if (xs >= len) return;
    // this warp's 256 elements: valid pairs before them inside the slice
    const int xe = min(len, xs + PK_DC_SLICE);
    const int xw0 = xs + wid * 256;
    uint32_t m[8];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int xw = xw0 + j * 32;
        m[j] = xw < xe ? pk_pair_word(s_vbc, xw, d) : 0u;
        if (xw < xe && xw + 32 > xe) m[j] &= (1u << (xe - xw)) - 1u;
        mine += __popc(m[j]);
    }
    __syncthreads();
    if (lane == 0) s_part[wid] = mine;
    __syncthreads();
    int base = before;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < wid) base += s_part[k];
    double* sc = scratch + (long long)d * pitch;
    const int32_t* row = band + (long long)d * pitch;
    int c[8]; double wa[8], wb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int x = xw0 + j * 32 + lane;
        const bool f = (m[j] >> lane) & 1u;
        c[j] = f ? __ldg(row + x) : 0;
        wa[j] = (f && balanced) ? __ldg(w + x) : 0.0;
        wb[j] = (f && balanced) ? __ldg(w + x + d) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if ((m[j] >> lane) & 1u) {
            double val;
            if (c[j] == 0) val = 0.0;
            else if (!balanced) val = (double)c[j];
            else {
                val = __dmul_rn(__dmul_rn(wa[j], wb[j]), (double)c[j]);            // pk_value
                if (!(pk_weight_tame(wa[j]) && pk_weight_tame(wb[j])) && !isfinite(val)) val = 0.0;
            }
            sc[base + __popc(m[j] & ((1u << lane) - 1u))] = val;
        }
        base += __popc(m[j]);
    }
}

// Sums: one CTA per distance over the compacted row (k_diag_compact). numpy's pairwise tree is built level by
// level in shared memory by ONE warp (warp barriers only: the table is a few hundred segments, block-wide
// barriers would cost more than the work), the leaf sums are taken by all threads, an octet of lanes per
// leaf, and the same warp combines them back up the tree. 256 threads and few registers: every distance's
// CTA is resident at once (the kernel is bound by the latency of its dependent steps, not by bandwidth).
template <int CAP>
__global__ void __launch_bounds__(PK_DS_THREADS) k_diag_sums(
    long long pitch, const double* __restrict__ scratch,
    double* __restrict__ out_sum, const long long* __restrict__ out_cnt, int32_t* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char ds_raw[];
    DiagSmem<CAP>& sm = *reinterpret_cast<DiagSmem<CAP>*>(ds_raw);
    __shared__ int s_L, s_cur, s_overflow;
    const int d = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* sc = scratch + (long long)d * pitch;
    const int nd = (int)out_cnt[d];
    // ---- leaf table, level by level (warp 0) ----
    if (wid == 0) {
        if (lane == 0) { sm.seg_s[0][0] = 0; sm.seg_m[0][0] = nd; sm.nseg[0] = 1; }
        __syncwarp();
        int cur = 0, L = 0;
        bool overflow = false;
        for (;; ++L) {
            const int ns = sm.nseg[L];
            const int nw = (ns + 31) / 32;
            int carry = 0;
            for (int w0 = 0; w0 < nw; ++w0) {               // split mask of this level, word by word, with its prefix
                const int i = w0 * 32 + lane;
                const bool sp = (i < ns) && sm.seg_m[cur][i] > 128;
                const unsigned bal = __ballot_sync(0xffffffffu, sp);
                if (lane == 0) { sm.split[L][w0] = bal; sm.wpre[L][w0] = (uint16_t)carry; }
                carry += __popc(bal);
            }
            const int ns_next = ns + carry;
            if (lane == 0) sm.nseg[L + 1] = ns_next;
            __syncwarp();
            if (ns_next == ns) break;                        // nothing split: level L holds the leaves
            if (ns_next > CAP || L + 1 >= PK_DS_LEVELS) { overflow = true; break; }
            for (int i = lane; i < ns; i += 32) {
                const uint32_t word = sm.split[L][i >> 5];
                const int pos = i + sm.wpre[L][i >> 5] + __popc(word & ((1u << (i & 31)) - 1u));
                const int s0 = sm.seg_s[cur][i], m = sm.seg_m[cur][i];
                if ((word >> (i & 31)) & 1u) {
                    int n2 = m / 2;
                    n2 -= n2 % 8;
                    sm.seg_s[cur ^ 1][pos] = s0;          sm.seg_m[cur ^ 1][pos] = n2;
                    sm.seg_s[cur ^ 1][pos + 1] = s0 + n2; sm.seg_m[cur ^ 1][pos + 1] = m - n2;
                } else {
                    sm.seg_s[cur ^ 1][pos] = s0; sm.seg_m[cur ^ 1][pos] = m;
                }
            }
            __syncwarp();
            cur ^= 1;
        }
        if (lane == 0) { s_L = L; s_cur = cur; s_overflow = overflow ? 1 : 0; }
    }
    __syncthreads();
    if (s_overflow) {                       // diagonal longer than the shared-memory tables: refuse loudly
        if (tid == 0) { atomicOr(&flags[2], 2); out_sum[d] = CUDART_NAN; }
        return;
    }
    const int L = s_L, cur = s_cur;
    // ---- leaf sums (8 strided accumulators + tail): an octet of lanes per leaf ----
    const int nl = sm.nseg[L];
    constexpr int OCT = PK_DS_THREADS / 8;
    for (int l0 = 0; l0 < nl; l0 += OCT) {                        // uniform trip count: shuffles inside
        const int l = l0 + (tid >> 3);
        const bool have = l < nl;
        const double r = pk_leaf_sum8(sc + (have ? sm.seg_s[cur][l] : 0), have ? sm.seg_m[cur][l] : 0, tid & 7);
        if (have && (tid & 7) == 0) sm.val[0][l] = r;
    }
    __syncthreads();
    // ---- combine back up: left + right wherever a segment was split (warp 0) ----
    if (wid == 0) {
        int vc = 0;
        for (int lv = L - 1; lv >= 0; --lv) {
            const int ns = sm.nseg[lv];
            for (int i = lane; i < ns; i += 32) {
                const uint32_t word = sm.split[lv][i >> 5];
                const int pos = i + sm.wpre[lv][i >> 5] + __popc(word & ((1u << (i & 31)) - 1u));
                double v = sm.val[vc][pos];
                if ((word >> (i & 31)) & 1u) v = __dadd_rn(v, sm.val[vc][pos + 1]);
                sm.val[vc ^ 1][i] = v;
            }
            __syncwarp();
            vc ^= 1;
        }
        if (lane == 0) out_sum[d] = sm.val[vc][0];
    }
}

// ---------------------------------------------------------------------------
// S3  expected curve on the device (utils.py:160-176): mean where more than 10
//     valid entries; IsotonicRegression(increasing=False, out_of_bounds='clip')
//     = PAVA on the reversed positive means (scipy, Busing 2022 Alg. 1), drop
//     interior points of constant runs (sklearn), clip the query, numpy.interp.
//     Same operations, same order as pk_host.cpp::pk_fit_expected_host, which the
//     CPU tests pin against scikit-learn. One CTA; only PAVA itself is sequential.
// ---------------------------------------------------------------------------
#define PK_FIT_MAX 1024
#define PK_FIT_THREADS 256

// ordered compaction helper: exclusive prefix of per-thread counts over the block
__device__ __forceinline__ int pk_block_excl_scan(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    int pre = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < PK_FIT_THREADS / 32; ++k) { const int t = s_warp[k]; if (k < wid) pre += t; tot += t; }
    __syncthreads();
    *total = tot;
    return pre + x - v;
}

__global__ void __launch_bounds__(PK_FIT_THREADS) k_fit_expected(
    const double* __restrict__ sum, const long long* __restrict__ cnt, int len, double* __restrict__ out_exp,
    double* __restrict__ out_bg, int32_t* __restrict__ flags) {
    __shared__ double s_e[PK_FIT_MAX];      // PAVA values (reversed order)
    __shared__ double s_w[PK_FIT_MAX];
    __shared__ double s_kx[PK_FIT_MAX], s_ky[PK_FIT_MAX];
    __shared__ int s_xs[PK_FIT_MAX], s_r[PK_FIT_MAX + 1];
    __shared__ int s_warp[PK_FIT_THREADS / 32];
    const int tid = threadIdx.x;
    constexpr int PER = PK_FIT_MAX / PK_FIT_THREADS;       // consecutive distances per thread
    // ---- means and ordered compaction of the positive ones ----
    double e[PER];
    int npos = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int d = tid * PER + j;
        e[j] = 0.0;
        if (d < len && cnt[d] > 10) e[j] = __ddiv_rn(sum[d], (double)cnt[d]);
        npos += e[j] > 0.0;
    }
    int n;
    int at = pk_block_excl_scan(npos, s_warp, &n);
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (e[j] > 0.0) { s_xs[at] = tid * PER + j; s_ky[at] = e[j]; ++at; }
    __syncthreads();
    if (n == 0) {
        if (tid == 0) atomicOr(&flags[2], 1);      // no positive mean: the reference raises here
        for (int d = tid; d < len; d += PK_FIT_THREADS) { out_exp[d] = CUDART_NAN; out_bg[d] = CUDART_NAN; }
        return;
    }
    for (int i = tid; i < n; i += PK_FIT_THREADS) { s_e[i] = s_ky[n - 1 - i]; s_w[i] = 1.0; }
    __syncthreads();
    // s_ky is dead until the knots: it holds RN(1 / k), k = 1..PK_FIT_MAX, during PAVA, which divides by
    // block weights (sums of unit weights: small integers)
    double* s_rcp = s_ky;
    for (int k = tid; k < PK_FIT_MAX; k += PK_FIT_THREADS) s_rcp[k] = __ddiv_rn(1.0, (double)(k + 1));
    __syncthreads();
    // ---- PAVA on the reversed values ----
    // scipy's loop (Busing 2022, Alg. 1) visits every point, but a point only costs anything where the
    // sequence is out of order. Blocks are kept in place, at the index of their last point -- value s_e[p],
    // weight s_w[p], first point s_r[p]; every point starts as its own block, set up by all threads -- and
    // s_nv[p] is the next index >= p whose raw predecessor is not smaller (found by a warp with ballots).
    // The sequential part then jumps from one violation to the next and performs exactly the reference's
    // additions, multiplications and divisions there, in the reference's order. Block weights are sums of
    // unit weights, i.e. exact integers <= n: sb / wb goes through a tabulated reciprocal with two residual
    // corrections (pk_div_r: bit-identical to IEEE division inside its guarded range, else __ddiv_rn).
    int* s_nv = reinterpret_cast<int*>(s_kx);               // [n + 1]; s_kx is not needed before the expansion
    __shared__ unsigned char s_end[PK_FIT_MAX];
    for (int i = tid; i < n; i += PK_FIT_THREADS) { s_r[i] = i; s_end[i] = 1; }
    if (tid < 32) {
        int carry = n;                                      // next violation at or after the chunk to the right
        for (int c0 = ((n - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
            const int p = c0 + tid;
            const bool f = p >= 1 && p < n && s_e[p - 1] >= s_e[p];
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            const unsigned at_or_after = bal >> tid;
            if (p <= n) s_nv[p] = at_or_after ? p + __ffs(at_or_after) - 1 : carry;
            if (bal) carry = c0 + __ffs(bal) - 1;
        }
        if (tid == 0) s_nv[n] = n;
    }
    __syncthreads();
    auto div_w = [&](double sb, double wb) -> double {
        const int k = (int)wb;
        if (pk_div_safe(sb) && k >= 1 && k <= PK_FIT_MAX && (double)k == wb) return pk_div_r(sb, wb, s_rcp[k - 1]);
        return __ddiv_rn(sb, wb);
    };
    if (tid == 0) {
        int i = s_nv[1];
        while (i < n) {
            const int q = i - 1;                            // the block that ends just before point i
            if (!(s_e[q] >= s_e[i])) { i = s_nv[i + 1]; continue; }      // in order: on to the next raw violation
            double sb = __dadd_rn(__dmul_rn(s_w[q], s_e[q]), __dmul_rn(1.0, s_e[i]));
            double wb = __dadd_rn(1.0, s_w[q]);
            double xb = div_w(sb, wb);
            int st = s_r[q];
            s_end[q] = 0;
            while (i < n - 1 && xb >= s_e[i + 1]) {
                s_end[i] = 0;
                i++;
                sb = __dadd_rn(sb, __dmul_rn(1.0, s_e[i]));
                wb = __dadd_rn(wb, 1.0);
                xb = div_w(sb, wb);
            }
            while (st > 0 && s_e[st - 1] >= xb) {
                const int q2 = st - 1;
                sb = __dadd_rn(sb, __dmul_rn(s_w[q2], s_e[q2]));
                wb = __dadd_rn(wb, s_w[q2]);
                xb = div_w(sb, wb);
                s_end[q2] = 0;
                st = s_r[q2];
            }
            s_e[i] = xb; s_w[i] = wb; s_r[i] = st;
            i++;
        }
    }
    __syncthreads();
    // ---- expand the blocks into the forward-ordered fit (s_kx as yf[]; s_nv is dead) ----
    {
        double xk[PER]; int lo[PER], hi[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int p = tid * PER + j;
            hi[j] = -1; lo[j] = 0; xk[j] = 0.0;
            if (p < n && s_end[p]) { hi[j] = p; lo[j] = s_r[p]; xk[j] = s_e[p]; }
        }
        __syncthreads();                                    // every read of s_nv's memory precedes the writes below
#pragma unroll
        for (int j = 0; j < PER; ++j)
            for (int i = lo[j]; i <= hi[j]; ++i) s_kx[n - 1 - i] = xk[j];
    }
    __syncthreads();
    // ---- knots: first, last, and every point that differs from a neighbour ----
    double yv[PER]; int xv[PER]; int nk = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int i = tid * PER + j;
        bool keep = false;
        if (i < n) {
            const double yi = s_kx[i];
            keep = (i == 0 || i == n - 1) || (yi != s_kx[i - 1]) || (yi != s_kx[i + 1]);
            yv[j] = yi; xv[j] = s_xs[i];
        }
        if (!keep) xv[j] = -1;
        nk += keep;
    }
    int m;
    int kat = pk_block_excl_scan(nk, s_warp, &m);      // (its barriers also order the reads of s_kx above)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (xv[j] >= 0) { s_e[kat] = (double)xv[j]; s_w[kat] = yv[j]; ++kat; }   // knots: x in s_e, y in s_w
    __syncthreads();
    const double xmin = (double)s_xs[0], xmax = (double)s_xs[n - 1];
    for (int d = tid; d < len; d += PK_FIT_THREADS) {
        const double T = fmin(fmax((double)d, xmin), xmax);
        double v;
        if (m == 1) {
            v = s_w[0];
        } else {
            int lo = 0, hi = m;              // upper_bound: first knot > T
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_e[mid] <= T) lo = mid + 1; else hi = mid;
            }
            int j = lo - 1;
            j = j < 0 ? 0 : (j > m - 1 ? m - 1 : j);
            if (j == m - 1 || s_e[j] == T) {
                v = s_w[j];
            } else {
                const double slope = __ddiv_rn(__dsub_rn(s_w[j + 1], s_w[j]), __dsub_rn(s_e[j + 1], s_e[j]));
                v = __dadd_rn(__dmul_rn(slope, __dsub_rn(T, s_e[j])), s_w[j]);
            }
        }
        out_exp[d] = v;
        out_bg[d] = v;
    }
}

// ---------------------------------------------------------------------------
// S4  Poisson candidate scan (scoreUtils.py:40-68).
//     candidate <=> count > 0 and mu = bg[d] / (w_x * w_y) satisfies 0 <= mu < crit[count].
//     Reference order is distance asc, row asc, and every candidate needs its rank in that order over
//     the whole chromosome (the 100,000-candidate batches of scoreUtils.py:104) and its position in the
//     list of this row tile. Two kernels over tiles of PK_CTILE band slots (distance-major):
//       k_cand_mark   evaluates every slot once (8 per thread and round, loads first), keeps one bit per
//                     slot and the tile's two counts;
//       k_cand_write  gets the counts of all earlier tiles by summing them directly (a few thousand values
//                     from L2 -- no scan kernel, no waiting on other CTAs) and expands its bits into the
//                     ordered candidate list; the last tile leaves the totals.
// ---------------------------------------------------------------------------
#define PK_CTILE 4096

// the exact test of a slot whose approximate comparison is too close to call (rare): kept out of line so that the
// sixteen unrolled slots of a thread do not each carry a division
__device__ __noinline__ bool pk_poisson_exact(double e, double p, double cr) {
    const double mu = __ddiv_rn(e, p);
    return (mu >= 0.0) && (mu < cr);
}

__global__ void __launch_bounds__(256) k_cand_mark(
    const int32_t* __restrict__ band, const double* __restrict__ w, const double* __restrict__ bg,
    int n, long long pitch, int balanced, int lower, const double* __restrict__ crit, int kmax,
    int row_begin, int row_end, int n_chunks, uint32_t* __restrict__ bits,
    uint2* __restrict__ counts, int32_t* __restrict__ flags) {
    const int chunk = blockIdx.x, di = blockIdx.y, d = lower + di;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int len = n - d;
    const double e = bg[d];
    const bool d_ok = (len > 0) && (e > 0.0);
    const int32_t* row = band + (long long)d * pitch;
    const long long tile = (long long)di * n_chunks + chunk;
    __shared__ int s_a[8], s_t[8];
    int tot_a = 0, tot_t = 0;
    bool over = false;                           // a count beyond the table of critical means
    // mu < crit[k]  <=>  e / p < crit[k], p = w_x w_y. The division is only needed when e and crit[k] * p are
    // within 2^-40 of each other: both roundings (of the product here, of the quotient there) are below
    // 2^-52 relative, so outside that band the comparison of e with crit[k] * p decides -- and decides the
    // same. e outside [1e-290, 1e290] (denormal products) always takes the division.
    const bool approx_ok = e > 1e-290 && e < 1e290;
    const double e_hi = __dmul_rn(e, 1.0 + 0x1p-40), e_lo = __dmul_rn(e, 1.0 - 0x1p-40);
    const bool whole = row_begin <= 0 && row_end >= n;
    uint32_t* tb = bits + tile * (PK_CTILE / 32) + wid;          // word q of the tile covers slots 32 q .. 32 q + 31
#pragma unroll
    for (int r = 0; r < PK_CTILE / 2048; ++r) {
        // counts and weights of 8 slots are loaded side by side (one memory round trip)
        int k[8];
        double wx[8], wy[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int x = chunk * PK_CTILE + (r * 8 + j) * 256 + tid;
            const bool in = d_ok && x < len;
            k[j] = in ? row[x] : 0;
            wx[j] = (balanced && in) ? w[x] : 1.0;
            wy[j] = (balanced && in) ? w[x + d] : 1.0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            bool c = false;
            if (k[j] > 0) {
                if (k[j] > kmax) over = true;
                else {
                    const double cr = crit[k[j]];
                    if (!balanced) c = (e >= 0.0) && (e < cr);
                    else {
                        const double p = __dmul_rn(wx[j], wy[j]);
                        const double t = __dmul_rn(cr, p);
                        if (approx_ok && t > e_hi) c = true;
                        else if (approx_ok && t < e_lo) c = false;
                        else c = pk_poisson_exact(e, p, cr);
                    }
                }
            }
            const unsigned ba = __ballot_sync(0xffffffffu, c);
            unsigned bt = ba;
            if (!whole) {
                const int x = chunk * PK_CTILE + (r * 8 + j) * 256 + tid;
                bt = __ballot_sync(0xffffffffu, c && x >= row_begin && x < row_end);
            }
            if (lane == 0) tb[(r * 8 + j) * 8] = ba;
            tot_a += __popc(ba);
            tot_t += __popc(bt);
        }
    }
    if (__any_sync(0xffffffffu, over) && lane == 0) atomicOr(&flags[0], 1);
    if (lane == 0) { s_a[wid] = tot_a; s_t[wid] = tot_t; }
    __syncthreads();
    if (tid == 0) {
        int a = 0, t = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { a += s_a[q]; t += s_t[q]; }
        counts[tile] = make_uint2((unsigned)a, (unsigned)t);
    }
}

__global__ void __launch_bounds__(256) k_cand_write(
    const uint32_t* __restrict__ bits, const uint2* __restrict__ counts, int lower, int row_begin, int row_end,
    int n_chunks, long long n_tiles, long long cap, int32_t* __restrict__ cx, int32_t* __restrict__ cd,
    int32_t* __restrict__ crank, long long* __restrict__ ncand, int32_t* __restrict__ flags) {
    const int chunk = blockIdx.x, di = blockIdx.y, d = lower + di;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long tile = (long long)di * n_chunks + chunk;
    const uint2 mine = counts[tile];
    const bool is_last = tile == n_tiles - 1;
    if (mine.x == 0 && !is_last) return;                       // no candidate in this tile
    // candidates in all earlier tiles (distance-major order): a direct sum
    unsigned long long pa = 0, pt = 0;
    for (long long i = tid; i < tile; i += 256) { const uint2 c = counts[i]; pa += c.x; pt += c.y; }
    __shared__ unsigned long long s_pa[8], s_pt[8];
    __shared__ uint32_t s_wa[PK_CTILE / 32], s_wt[PK_CTILE / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { pa += __shfl_xor_sync(0xffffffffu, pa, o); pt += __shfl_xor_sync(0xffffffffu, pt, o); }
    if (lane == 0) { s_pa[wid] = pa; s_pt[wid] = pt; }
    // per-word candidate counts of this tile (all rows / rows of the row tile), then their exclusive prefix
    constexpr int NWORD = PK_CTILE / 32;                       // 128 words, one per thread of the first four warps
    uint32_t word = 0, inr = 0;
    if (tid < NWORD) {
        word = bits[tile * NWORD + tid];
        // word q = (round j) * 8 + warp covers slots x = chunk * PK_CTILE + j * 256 + warp * 32 + bit
        const int xw = chunk * PK_CTILE + (tid >> 3) * 256 + (tid & 7) * 32;
        const int lo = max(row_begin - xw, 0), hi = min(row_end - xw, 32);
        if (hi > lo) inr = (hi - lo >= 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
        s_wa[tid] = __popc(word);
        s_wt[tid] = __popc(word & inr);
    }
    __syncthreads();
    pa = 0; pt = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { pa += s_pa[q]; pt += s_pt[q]; }
    if (is_last && tid == 0) { ncand[0] = (long long)(pt + mine.y); ncand[1] = (long long)(pa + mine.x); }
    if (mine.x == 0) return;
    if (tid < NWORD) {
        // exclusive prefix over the 128 words: slot order is word order
        unsigned ea = 0, et = 0;
        for (int q = 0; q < tid; ++q) { ea += s_wa[q]; et += s_wt[q]; }
        const int j = tid >> 3, wq = tid & 7;
        uint32_t rem = word & inr;
        while (rem) {
            const int bit = __ffs(rem) - 1;
            rem &= rem - 1;
            const unsigned lm = (1u << bit) - 1u;
            const long long rt = (long long)pt + et + __popc(word & inr & lm);
            if (rt < cap) {
                cx[rt] = chunk * PK_CTILE + j * 256 + wq * 32 + bit;
                cd[rt] = d;
                crank[rt] = (int32_t)(pa + ea + __popc(word & lm));
            } else {
                atomicOr(&flags[3], 2);                        // candidate buffer too small
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
template <typename TY, typename TC, bool DELTA>
static int launch_band_csr_t(pk_chrom* c, const long long* rowptr, const void* b2, const void* cnt) {
    // Rows per CTA: 32, or a few more when that lets every CTA be resident at once. The band
    // barely fits the GPU's shared memory, so with 32 rows a chr1-scale chromosome needs 779
    // CTAs for 740 slots -- a second wave for 5 % of the work.
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    auto tile_bytes = [&](int R) { return (size_t)c->ND * (R + 1) * sizeof(int32_t); };
    auto slots = [&](int R) { return (long long)sms * std::min<long long>(8, (long long)((227 * 1024) / (tile_bytes(R) + 1024))); };
    int R = 32;
    if ((c->n + 31) / 32 > slots(32))
        for (int r = 34; r <= 64; r += 2)
            if ((c->n + r - 1) / r <= slots(r)) { R = r; break; }
    const size_t smem = tile_bytes(R);
    if (smem > 200 * 1024) { pk_set_error("band build: %d diagonals do not fit a shared-memory tile", c->ND); return PK_EUNSUPPORTED; }
    PK_OPT_IN_SMEM((k_band_csr<TY, TC, DELTA>), smem, c->device);
    const unsigned grid = (unsigned)((c->n + R - 1) / R);
    k_band_csr<TY, TC, DELTA><<<grid, 256, smem, c->stream>>>(rowptr, (const TY*)b2, (const TC*)cnt, c->d_w, c->n, c->ND, c->pitch,
                                                              c->balanced, R, c->d_band, c->d_valid, c->d_flags);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

static PkRowsView rows_view(const pk_chrom* c) {
    const unsigned char* base = c->d_blob;
    const long long* h = c->rows_hdr;
    PkRowsView v;
    v.nd_enc = (int)h[2]; v.W = (int)h[3]; v.n_esc = h[5];
    v.bits = reinterpret_cast<const uint32_t*>(base + h[7]);
    v.cnt_off = reinterpret_cast<const uint32_t*>(base + h[8]);
    v.cnt8 = base + h[9];
    v.esc = reinterpret_cast<const int32_t*>(base + h[10]);
    v.far_off = reinterpret_cast<const long long*>(base + h[11]);
    v.far_b2 = reinterpret_cast<const int32_t*>(base + h[12]);
    v.far_cnt = reinterpret_cast<const int32_t*>(base + h[13]);
    return v;
}

// rows per CTA of the tiled band builds: 32, or a few more when that lets every CTA be resident at once
static int band_rows_per_cta(const pk_chrom* c) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    auto tile_bytes = [&](int R) { return (size_t)c->ND * (R + 1) * sizeof(int32_t); };
    auto slots = [&](int R) { return (long long)sms * std::min<long long>(8, (long long)((227 * 1024) / (tile_bytes(R) + 1024))); };
    int R = 32;
    if ((c->n + 31) / 32 > slots(32))
        for (int r = 34; r <= 64; r += 2)
            if ((c->n + r - 1) / r <= slots(r)) { R = r; break; }
    return R;
}

int pk_launch_band_rows(pk_chrom* c) {
    const PkRowsView v = rows_view(c);
    const int R = band_rows_per_cta(c);
    const size_t smem = (size_t)c->ND * (R + 1) * sizeof(int32_t);
    if (smem > 200 * 1024) { pk_set_error("band build: %d diagonals do not fit a shared-memory tile", c->ND); return PK_EUNSUPPORTED; }
    PK_OPT_IN_SMEM(k_band_rows, smem, c->device);
    k_band_rows<<<(unsigned)((c->n + R - 1) / R), 256, smem, c->stream>>>(v, c->d_w, c->n, c->ND, c->pitch, c->balanced, R,
                                                                       c->d_band, c->d_valid, c->d_flags);
    PK_CUDA(cudaGetLastError());
    if (v.n_esc > 0) {
        k_band_escapes<<<(unsigned)((v.n_esc + 255) / 256), 256, 0, c->stream>>>(v, c->d_w, c->n, c->ND, c->pitch, c->balanced,
                                                                              c->d_band, c->d_valid, c->d_flags);
        PK_CUDA(cudaGetLastError());
    }
    return PK_OK;
}

// enc 0: int32 bin2 + int32 count; enc 1: uint16 (bin2 - bin1) + uint16 count
int pk_launch_band_csr(pk_chrom* c, const long long* rowptr, const void* b2, const void* cnt, int enc) {
    if (enc == 1) return launch_band_csr_t<uint16_t, uint16_t, true>(c, rowptr, b2, cnt);
    return launch_band_csr_t<int32_t, int32_t, false>(c, rowptr, b2, cnt);
}

// ---------------------------------------------------------------------------
// depth (calculate_depth.py:25-28): sum of the raw counts with bin2 - bin1 >= min_dis over the
// uploaded pixel columns. A warp per row for row-indexed columns, a thread per pixel for COO.
// ---------------------------------------------------------------------------
template <typename TY, typename TC, bool DELTA>
__global__ void __launch_bounds__(256) k_depth_rows(const long long* __restrict__ rowptr, const TY* __restrict__ b2,
                                                    const TC* __restrict__ cnt, int n, int min_dis,
                                                    unsigned long long* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * 256) >> 5;
    unsigned long long sum = 0;
    for (int x = (blockIdx.x * 256 + threadIdx.x) >> 5; x < n; x += warps) {
        const long long p0 = rowptr[x], p1 = rowptr[x + 1];
        for (long long p = p0 + lane; p < p1; p += 32) {
            const int d = DELTA ? (int)b2[p] : (int)b2[p] - x;
            const long long c = (long long)cnt[p];
            if (d >= min_dis && c > 0) sum += (unsigned long long)c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0 && sum) atomicAdd(total, sum);
}

__global__ void __launch_bounds__(256) k_depth_coo(const int32_t* __restrict__ b1, const int32_t* __restrict__ b2,
                                                   const int32_t* __restrict__ cnt, long long nnz, int min_dis,
                                                   unsigned long long* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * 256;
    unsigned long long sum = 0;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < nnz; p += stride) {
        const int d = b2[p] - b1[p];
        if (d >= min_dis && cnt[p] > 0) sum += (unsigned long long)cnt[p];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0 && sum) atomicAdd(total, sum);
}

int pk_launch_depth(pk_chrom* c, int32_t min_dis, unsigned long long* d_total) {
    const unsigned grid = 148 * 4;
    if (c->up_kind == 4) {
        k_depth_packed<<<grid, 256, 0, c->stream>>>(rows_view(c), c->n, min_dis, d_total);
        PK_CUDA(cudaGetLastError());
        return PK_OK;
    }
    if (c->up_nnz == 0) return PK_OK;
    if (c->up_kind == 1)
        k_depth_coo<<<grid, 256, 0, c->stream>>>((const int32_t*)c->up_b1, (const int32_t*)c->up_b2, (const int32_t*)c->up_cnt,
                                                 c->up_nnz, min_dis, d_total);
    else if (c->up_kind == 3)
        k_depth_rows<uint16_t, uint16_t, true><<<grid, 256, 0, c->stream>>>(c->up_rowptr, (const uint16_t*)c->up_b2,
                                                                            (const uint16_t*)c->up_cnt, c->n, min_dis, d_total);
    else
        k_depth_rows<int32_t, int32_t, false><<<grid, 256, 0, c->stream>>>(c->up_rowptr, (const int32_t*)c->up_b2,
                                                                           (const int32_t*)c->up_cnt, c->n, min_dis, d_total);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int pk_launch_rowptr(pk_chrom* c, const int32_t* b1, const int32_t* b2, int64_t nnz, long long* rowptr) {
    if (nnz > 0) {
        k_check_sorted<<<(unsigned)((nnz + 255) / 256), 256, 0, c->stream>>>(b1, b2, nnz, c->d_flags);
        PK_CUDA(cudaGetLastError());
    }
    k_rowptr<<<(unsigned)((c->n + 1 + 255) / 256), 256, 0, c->stream>>>(b1, nnz, c->n, rowptr);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

template <int CAP>
static int launch_diag_sums_t(pk_chrom* c) {
    const size_t smem = sizeof(DiagSmem<CAP>);
    PK_OPT_IN_SMEM(k_diag_sums<CAP>, smem, c->device);
    k_diag_sums<CAP><<<c->ND, PK_DS_THREADS, smem, c->stream>>>(c->pitch, c->d_scratch, c->d_diag_sum, c->d_diag_cnt, c->d_flags);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int pk_launch_diag_sums(pk_chrom* c) {
    const int n_words = (c->n + 31) / 32;
    k_valid_bits<<<std::min((n_words + 7) / 8, 148 * 4), 256, 0, c->stream>>>(c->d_valid, c->d_w, c->balanced, c->n, c->d_vbits, n_words,
                                                                c->d_flags);
    PK_CUDA(cudaGetLastError());
    const size_t smem_c = ((size_t)n_words + 2) * 4;
    if (smem_c > 200 * 1024) { pk_set_error("chromosome of %d bins: valid mask does not fit shared memory", c->n); return PK_EUNSUPPORTED; }
    PK_OPT_IN_SMEM(k_diag_compact, smem_c, c->device);
    dim3 grid((unsigned)((c->n + PK_DC_SLICE - 1) / PK_DC_SLICE), (unsigned)c->ND);
    k_diag_compact<<<grid, 256, smem_c, c->stream>>>(c->d_band, c->d_w, c->d_vbits, n_words, c->n, c->pitch, c->balanced,
                                                    c->d_scratch, c->d_diag_cnt);
    PK_CUDA(cudaGetLastError());
    // leaves of numpy's tree hold at least 57 elements
    if (c->n <= 57 * 1024) return launch_diag_sums_t<1024>(c);
    return launch_diag_sums_t<4096>(c);
}

bool pk_fit_on_device_supported(int len) { return len <= PK_FIT_MAX; }

int pk_launch_fit_expected(pk_chrom* c) {
    k_fit_expected<<<1, PK_FIT_THREADS, 0, c->stream>>>(c->d_diag_sum, c->d_diag_cnt, c->ND, c->d_exp, c->d_bg, c->d_flags);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int pk_launch_candidates(pk_chrom* c, const double* d_crit, int kmax) {
    const int nd = c->upper - c->lower + 1;
    if (nd <= 0) return PK_OK;
    const long long m = (long long)nd * c->n_chunks;
    dim3 grid(c->n_chunks, nd);
    uint2* counts = reinterpret_cast<uint2*>(c->d_cstate);
    k_cand_mark<<<grid, 256, 0, c->stream>>>(c->d_band, c->use_wp ? c->d_wp : c->d_w, c->d_bg, c->n, c->pitch, c->balanced, c->lower, d_crit, kmax,
                                            c->row_begin, c->row_end, c->n_chunks, c->d_bits, counts, c->d_flags);
    PK_CUDA(cudaGetLastError());
    k_cand_write<<<grid, 256, 0, c->stream>>>(c->d_bits, counts, c->lower, c->row_begin, c->row_end, c->n_chunks, m, c->cand_cap,
                                             c->d_cx, c->d_cd, c->d_crank, c->d_ncand, c->d_flags);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}
