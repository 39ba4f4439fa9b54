// peakachu_b200: fused window-features + forest kernel (the dominant stage).
//
// One persistent CTA of P threads per SM slot. It repeats:
//   phase A  (scoreUtils.py:70-93)  fill a shared-memory feature buffer with up to P
//            windows that pass the reference's filters. A warp works on two candidates
//            at a time, one per 16-lane half; lane h of a half owns column h of the
//            (2W+1)^2 window for the vertical Gaussian pass and row h for the horizontal
//            pass, so both passes run in registers with one shared-memory transpose.
//   phase B  (scoreUtils.py:109)    one pixel per thread walks the forest. Trees are
//            staged group by group into two shared-memory buffers with TMA bulk copies
//            (cp.async.bulk + mbarrier), so node fetches are LDS instead of divergent
//            global loads; four trees are walked at once per thread for ILP, leaf
//            values are added in estimator order in float64.
// Features never leave the SM: HBM traffic is the band cells of the windows, the
// candidate list and one (keep, prob) pair per candidate.
#include "pk_common.cuh"
#include "pk_device.cuh"

#include <algorithm>

struct FusedParams {
    const int32_t* band; const double* w; const double* expv;
    int n; long long pitch; int balanced; int ND;
    const int32_t* cx; const int32_t* cd; const int32_t* crank; long long n_cand;
    const uint2* nodes; const uint32_t* roots; const uint8_t* depth; const int4* groups;
    int n_groups; int n_trees;
    uint8_t* keep; double* prob; int32_t* batch_win; unsigned long long* counters;
    unsigned long long* next;      // global work counter (candidates handed out)
};

// ---- mbarrier / bulk-copy PTX ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int W, int P>
struct FusedSmem {
    static constexpr int S = 2 * W + 1, F = S * S, NW = P / 32;
    static constexpr size_t fea_bytes = (size_t)P * F * 4;
    static constexpr size_t node_bytes = 2 * (size_t)PK_TREE_BUF_NODES * 8;
    static constexpr size_t scratch_bytes = (size_t)NW * 2 * F * (8 + 4);
    static size_t total(int ND) {
        return node_bytes + fea_bytes + (size_t)((ND + 1) & ~1) * 8 + scratch_bytes + (size_t)P * 4 + (size_t)F * 2 + 64;
    }
};

template <int W, int P>
__global__ void __launch_bounds__(P, (P <= 128 ? 2 : 1)) k_score_fused(const FusedParams prm) {
    constexpr int S = 2 * W + 1, F = S * S, NW = P / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: node buffers (16 B aligned) | features | exp | per-warp scratch | slot->candidate | lut | barriers
    uint2* s_nodes = reinterpret_cast<uint2*>(smem_raw);
    float* s_fea = reinterpret_cast<float*>(smem_raw + FusedSmem<W, P>::node_bytes);
    double* s_exp = reinterpret_cast<double*>(smem_raw + FusedSmem<W, P>::node_bytes + FusedSmem<W, P>::fea_bytes);
    const int ND = prm.ND, NDp = (ND + 1) & ~1;
    double* s_V = s_exp + NDp;                                    // [NW][2][F] float64
    int32_t* s_C = reinterpret_cast<int32_t*>(s_V + (size_t)NW * 2 * F);   // [NW][2][F] int32
    int32_t* s_idx = s_C + (size_t)NW * 2 * F;                    // [P]
    uint16_t* s_lut = reinterpret_cast<uint16_t*>(s_idx + P);     // [F] cell order, diagonal-major
    uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_lut + F) + 15) & ~(uintptr_t)15);
    __shared__ int s_nkept, s_take, s_done;
    __shared__ long long s_start;

    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int half = lane >> 4, h = lane & 15;
    const int G = prm.n_groups;
    const bool resident = (G <= 2);

    // ---- one-time setup -----------------------------------------------------
    for (int i = tid; i < ND; i += P) s_exp[i] = prm.expv[i];
    if (tid == 0) {
        // cells ordered by window diagonal (b - a), then along it: contiguous in the band
        int k = 0;
        for (int df = -(S - 1); df <= S - 1; ++df)
            for (int a = 0; a < S; ++a) {
                int b = a + df;
                if (b >= 0 && b < S) s_lut[k++] = (uint16_t)((a << 8) | b);
            }
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        s_done = 0;
    }
    __syncthreads();
    uint32_t issued = 0, consumed = 0;     // stream positions of tree-group loads (uniform across threads)
    auto issue = [&](uint32_t pos) {
        if (tid == 0) {
            const int4 g = prm.groups[pos % G];
            const uint32_t bytes = (uint32_t)g.w * 8u;
            uint64_t* bar = &s_bar[pos & 1];
            mbar_expect_tx(bar, bytes);
            bulk_g2s(s_nodes + (size_t)(pos & 1) * PK_TREE_BUF_NODES, prm.nodes + g.z, bytes, bar);
        }
    };
    issue(0); issued = 1;
    if (G > 1 || !resident) { issue(1); issued = 2; }
    bool first_batch = true;

    double* myV = s_V + (size_t)(wib * 2 + half) * F;
    int32_t* myC = s_C + (size_t)(wib * 2 + half) * F;

    for (;;) {
        // ================= phase A: features =================
        if (tid == 0) s_nkept = 0;
        __syncthreads();
        for (;;) {
            if (tid == 0) {
                int free_slots = P - s_nkept;
                long long st = (long long)atomicAdd(prm.next, (unsigned long long)free_slots);
                long long rem = prm.n_cand - st;
                s_start = st;
                s_take = rem <= 0 ? 0 : (int)(rem < free_slots ? rem : free_slots);
                if (rem <= free_slots) s_done = 1;
            }
            __syncthreads();
            const int take = s_take;
            const long long start = s_start;
            for (int j0 = wib * 2; j0 < take; j0 += NW * 2) {
                // ---- two candidates per warp, one per half ----
                const int j = j0 + half;
                const bool have = j < take;
                const long long ci = start + j;
                int x = 0, d = 0;
                if (have) { x = prm.cx[ci]; d = prm.cd[ci]; }
                const int y = x + d;
                bool ok = have && (x - W >= 0) && (y + W + 1 <= prm.n);        // scoreUtils.py:75
                // gather both windows with all 32 lanes, cells in diagonal-major order
                {
                    const int x0 = __shfl_sync(0xffffffffu, x, 0), d0 = __shfl_sync(0xffffffffu, d, 0);
                    const int x1 = __shfl_sync(0xffffffffu, x, 16), d1 = __shfl_sync(0xffffffffu, d, 16);
                    const bool ok0 = __shfl_sync(0xffffffffu, (int)ok, 0), ok1 = __shfl_sync(0xffffffffu, (int)ok, 16);
                    int32_t* C0 = s_C + (size_t)(wib * 2) * F;
#pragma unroll 2
                    for (int idx = lane; idx < 2 * F; idx += 32) {
                        const int k = idx >= F;
                        const int cell = s_lut[idx - k * F];
                        const int a = cell >> 8, b = cell & 255;
                        const int xx = k ? x1 : x0, dd0 = k ? d1 : d0;
                        if (k ? ok1 : ok0) {
                            const int r = xx - W + a, c = xx + dd0 - W + b;
                            const int dd = c - r, ad = dd < 0 ? -dd : dd, lo = dd < 0 ? c : r;
                            int cnt = 0;
                            if (ad < ND - 1) cnt = __ldg(prm.band + (long long)ad * prm.pitch + lo);   // scoreUtils.py:31
                            C0[(size_t)k * F + a * S + b] = cnt;
                        }
                    }
                }
                __syncwarp();
                const bool act = ok && (h < S);
                double v[S];
                int nz = 0;
                if (act) {
                    const double wc = prm.balanced ? prm.w[y - W + h] : 0.0;
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        const int cnt = myC[a * S + h];
                        const double wr = prm.balanced ? __ldg(prm.w + x - W + a) : 0.0;
                        v[a] = pk_value(cnt, wr, wc, prm.balanced);
                        nz += (v[a] != 0.0);
                        myV[a * S + h] = v[a];
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
                __syncwarp();
                if (ok) {
                    if ((double)nz < (double)F * 0.1) ok = false;              // utils.py:225
                }
                if (ok) {
                    double s = 0.0;                                            // utils.py:228 (numba order)
#pragma unroll
                    for (int a = 0; a < W; ++a)
#pragma unroll
                        for (int b = 0; b < W; ++b) s = __dadd_rn(s, myV[a * S + b]);
                    const double ll = __ddiv_rn(s, (double)(W * W));
                    ok = (ll > 0.0) && (__ddiv_rn(myV[W * S + W], ll) > 0.1);  // utils.py:229-232
                }
                __syncwarp();                                                  // V is overwritten below
                const bool kept = ok;                                          // uniform within the half
                const bool actk = kept && (h < S);
                int slot = 0;
                if (kept && h == 0) slot = atomicAdd(&s_nkept, 1);
                slot = __shfl_sync(0xffffffffu, slot, half * 16);
                double mn = CUDART_INF, mx = -CUDART_INF;
                bool has_nan = false;
                double g[S];
                if (actk) {
                    // distance normalisation (utils.py:187-200) + vertical pass in registers (column h)
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        int dd = d + h - a;
                        dd = dd < 0 ? -dd : dd;
                        v[a] = __ddiv_rn(v[a], s_exp[dd]);
                    }
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        double t = __dmul_rn(v[a], PK_GK[4]);
#pragma unroll
                        for (int jj = 4; jj >= 1; --jj)
                            t = __dadd_rn(t, __dmul_rn(__dadd_rn(v[pk_reflect(a - jj, S)], v[pk_reflect(a + jj, S)]), PK_GK[4 - jj]));
                        myV[a * S + h] = t;
                    }
                }
                __syncwarp();
                if (actk) {
                    // horizontal pass in registers (row h)
                    double t[S];
#pragma unroll
                    for (int b = 0; b < S; ++b) t[b] = myV[h * S + b];
#pragma unroll
                    for (int b = 0; b < S; ++b) {
                        double q = __dmul_rn(t[b], PK_GK[4]);
#pragma unroll
                        for (int jj = 4; jj >= 1; --jj)
                            q = __dadd_rn(q, __dmul_rn(__dadd_rn(t[pk_reflect(b - jj, S)], t[pk_reflect(b + jj, S)]), PK_GK[4 - jj]));
                        g[b] = q;
                        has_nan |= isnan(q);
                        mn = fmin(mn, q); mx = fmax(mx, q);
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    has_nan |= (bool)__shfl_xor_sync(0xffffffffu, (int)has_nan, o);
                }
                if (actk) {
                    if (has_nan) { mn = CUDART_NAN; mx = CUDART_NAN; }          // numba min/max propagate NaN
                    const double range = __dsub_rn(mx, mn);
                    float* frow = s_fea + (size_t)slot * F + h * S;
#pragma unroll
                    for (int b = 0; b < S; ++b)
                        frow[b] = __double2float_rn(__ddiv_rn(__dsub_rn(g[b], mn), range));   // utils.py:207
                    if (h == 0) {
                        s_idx[slot] = (int)ci;
                        prm.keep[ci] = 1;
                        atomicAdd(&prm.batch_win[prm.crank[ci] / PK_BATCH], 1);
                    }
                }
                __syncwarp();
            }
            __syncthreads();
            const int nk_now = s_nkept, done_now = s_done;
            __syncthreads();                       // everyone has read them before thread 0 grabs again
            if (done_now || nk_now > P - P / 8) break;
        }
        const int nkept = s_nkept;
        const bool last = s_done != 0;

        // ================= phase B: forest =================
        if (nkept > 0) {
            const bool mine = tid < nkept;
            const float* xrow = s_fea + (size_t)tid * F;
            double acc = 0.0;
            for (int gi = 0; gi < G; ++gi) {
                const uint32_t pos = resident ? (uint32_t)gi : consumed;
                if (!resident || first_batch) mbar_wait(&s_bar[pos & 1], (pos >> 1) & 1);
                const int4 grp = prm.groups[gi];
                const uint2* buf = s_nodes + (size_t)(pos & 1) * PK_TREE_BUF_NODES;
                const uint32_t gbase = (uint32_t)grp.z, staged = (uint32_t)grp.w;
                if (mine) {
                    for (int t = grp.x; t < grp.x + grp.y; t += 4) {
                        uint32_t p[4]; bool run[4]; int maxd = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const bool ex = (t + k) < grp.x + grp.y;
                            const uint32_t root = ex ? prm.roots[t + k] : 0x80000000u;
                            p[k] = root & 0x7fffffffu;
                            run[k] = ex && !(root >> 31);
                            const int dp = ex ? (int)prm.depth[t + k] : 0;
                            maxd = dp > maxd ? dp : maxd;
                        }
                        for (int lvl = 0; lvl < maxd; ++lvl) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (run[k]) {
                                    const uint32_t off = p[k] - gbase;
                                    const uint2 nd = off < staged ? buf[off] : __ldg(prm.nodes + p[k]);
                                    const float xv = xrow[nd.y & ((1u << PK_FEAT_BITS) - 1u)];
                                    const bool left = isnan(xv) ? ((nd.y >> 10) & 1u) : (xv <= __uint_as_float(nd.x));
                                    const bool leaf = left ? ((nd.y >> 11) & 1u) : ((nd.y >> 12) & 1u);
                                    p[k] = left ? p[k] + 1u : p[k] + (nd.y >> 13);
                                    run[k] = !leaf;
                                }
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if ((t + k) < grp.x + grp.y) {
                                const uint32_t off = p[k] - gbase;
                                const double lv = off < staged ? reinterpret_cast<const double*>(buf)[off]
                                                               : __ldg(reinterpret_cast<const double*>(prm.nodes) + p[k]);
                                acc = __dadd_rn(acc, lv);                      // estimator order
                            }
                        }
                    }
                }
                if (!resident) {
                    __syncthreads();                 // everyone is done with this buffer
                    ++consumed;
                    issue(issued); ++issued;         // refill it with the group two positions ahead
                }
            }
            if (mine) prm.prob[s_idx[tid]] = __ddiv_rn(acc, (double)prm.n_trees);
            if (tid == 0) atomicAdd(&prm.counters[1], (unsigned long long)nkept);
            first_batch = false;
        }
        if (last) break;
        __syncthreads();
    }
    // drain outstanding bulk copies before the CTA's shared memory is released
    if (!resident) {
        for (uint32_t pos = consumed; pos < issued; ++pos) mbar_wait(&s_bar[pos & 1], (pos >> 1) & 1);
    } else if (first_batch) {
        for (uint32_t pos = 0; pos < issued; ++pos) mbar_wait(&s_bar[pos & 1], 0);
    }
}

template <int W, int P>
static int launch_fused_t(const FusedParams& prm, int ND, int sm_count, cudaStream_t stream) {
    const size_t smem = FusedSmem<W, P>::total(ND);
    if (smem > 227 * 1024 - 256) { pk_set_error("fused kernel: %zu bytes of shared memory needed", smem); return PK_EUNSUPPORTED; }
    static size_t attr_set = 0;      // largest dynamic size opted into so far
    if (smem > attr_set) {
        PK_CUDA(cudaFuncSetAttribute(k_score_fused<W, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(P <= 128 ? 2 : 1, (227 * 1024) / smem));
    long long want = (prm.n_cand + P - 1) / P;
    unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * per_sm));
    k_score_fused<W, P><<<grid, P, smem, stream>>>(prm);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

// returns PK_EUNSUPPORTED (without touching the error string) when no fused variant fits
int pk_launch_fused(pk_chrom* c, const pk_forest* f, int variant) {
    if (c->n_cand == 0) return PK_OK;
    FusedParams prm;
    prm.band = c->d_band; prm.w = c->d_w; prm.expv = c->d_exp;
    prm.n = c->n; prm.pitch = c->pitch; prm.balanced = c->balanced; prm.ND = c->ND;
    prm.cx = c->d_cx; prm.cd = c->d_cd; prm.crank = c->d_crank; prm.n_cand = c->n_cand;
    prm.nodes = f->d_nodes; prm.roots = f->d_root; prm.depth = f->d_depth; prm.groups = f->d_groups;
    prm.n_groups = f->n_groups; prm.n_trees = f->n_trees;
    prm.keep = c->d_keep; prm.prob = c->d_prob; prm.batch_win = c->d_batch_win; prm.counters = c->d_counters;
    prm.next = c->d_counters + 2;
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, c->device);
    if (c->w == 5) return variant == 1 ? launch_fused_t<5, 256>(prm, c->ND, sm, c->stream)
                                       : launch_fused_t<5, 128>(prm, c->ND, sm, c->stream);
    if (c->w == 7) return variant == 1 ? launch_fused_t<7, 128>(prm, c->ND, sm, c->stream)
                                       : launch_fused_t<7, 64>(prm, c->ND, sm, c->stream);
    return PK_EUNSUPPORTED;
}

bool pk_fused_supported(int w) { return w == 5 || w == 7; }
