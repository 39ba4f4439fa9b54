// peakachu_b200: fused window-features + forest kernel (the dominant stage).
//
// One persistent CTA of NTH threads per SM, scoring up to P pixels per batch. It repeats:
//   phase A  (scoreUtils.py:70-93, utils.py:180-237)  fill the shared-memory feature buffer
//            (float32 [P][F]) with windows that pass the reference's filters. The CTA's warps form
//            NG independent groups (own staging buffer, own named barrier); a group takes PB
//            candidates at a time and runs five steps over the take: A1 gather + balance (a warp
//            owns window pairs, its 32 lanes share the 2*(2W+1)^2 band cells in band-contiguous
//            order, every load of the take issued before the first use), A2 the reference's
//            filters (one thread per window), A3 distance normalisation + vertical Gaussian pass
//            (one thread per window column), A4 horizontal pass + row extrema (one thread per
//            window row), A5 min-max scaling to float32 (a warp per window).
//   phase B  (scoreUtils.py:109)    TPP threads per pixel walk the forest. Trees are staged
//            group by group into two shared-memory buffers with TMA bulk copies
//            (cp.async.bulk + mbarrier) -- the same memory that held the float64 windows during
//            phase A -- so node fetches are LDS instead of divergent global loads. Each thread
//            walks CH trees at once, branch-free; leaf values are added in estimator order in
//            float64 (with TPP = 2 the second thread hands its leaf values over through shared
//            memory). Pixels that can no longer exceed --minimum-prob stop walking trees and the
//            survivors are re-packed onto the low threads.
// Features never leave the SM: HBM traffic is the band cells of the windows, the
// candidate list and one (keep, prob) pair per candidate.
#include "pk_common.cuh"
#include "pk_device.cuh"

#include <algorithm>
#include <cstring>

struct FusedParams {
    const int32_t* band; const double* w; const double* expv;
    int n; long long pitch; int balanced; int ND;
    const int32_t* cx; const int32_t* cd; const int32_t* crank;
    const long long* ncand_dev; long long cand_cap;   // candidate count lives on the device
    const uint2* nodes; const uint32_t* roots; const uint8_t* depth; const int4* groups;
    const uint8_t* rootfeat;       // child-feature encoding only (template parameter CF)
    const uint2* nodes_classic;    // pk_forest.d_nodes: missing_go_left of the child-feature encoding's NaN walk
    int n_groups; int n_trees;
    uint8_t* keep; double* prob; int32_t* batch_win; unsigned long long* counters;
    unsigned long long* next;      // global work counter (candidates handed out)
    const int32_t* flags;          // handle's device flags
    double thre;                   // --minimum-prob: pixels that can no longer exceed it stop walking trees
    float* fea_tap;                // parity tap (pk_chrom_fused_features): [n_cand][F] copy of the shared-memory feature rows, else NULL
    const int32_t* band2; long long P2;   // TM kernels: row-major band copy (pk_common.cuh) ...
    alignas(64) CUtensorMap tmap;         // ... and the skewed tensor map whose boxes are windows
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

// one window of the row-major band as a 2-D box (SASS UTMALDG.2D); c0 = first column (a multiple of 4), c1 = first row
__device__ __forceinline__ void tma_box_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

template <int W, int P, int TPP, int TBN, int CH, int NTH, int NGR = 2, int XSTAGE = 0, int CF = 0, int TM = 0>
struct FusedCfg {
    static constexpr int S = 2 * W + 1, F = S * S, NT = NTH, NW = NT / 32;      // NTH >= P*TPP: extra warps only build features
    static_assert(NTH >= P * TPP && NTH % 32 == 0, "thread count");
    static constexpr int NS = (2 * F + 31) / 32;              // gather slots per lane for a window pair
    static constexpr int NM = (F + 31) / 32;                  // cells per lane for one window
    static constexpr int CHUNK = CH * TPP;                    // trees walked per pixel per pass
    static constexpr size_t node_bytes = 2 * (size_t)TBN * 8;
    // phase B hand-over area: leaf values of the second sub-thread, running sums, list of pixels still walking
    static constexpr size_t hand_bytes = ((((TPP == 2 ? 2 * CH : 0) + 1) * (size_t)P * 8 + 4 * (size_t)P) + 15) & ~(size_t)15;
    // The tree buffers and the hand-over area are dead during phase A: the float64 windows are staged there.
    static constexpr size_t stage_bytes = node_bytes + hand_bytes + XSTAGE;   // XSTAGE: extra staging when shared memory is left
    // NG groups of warps build features independently (own staging buffer, own named barrier), so one
    // group's band gather overlaps the other's arithmetic
    static constexpr int NG = NGR, NTG = NT / NG, NWG = NW / NG;
    static_assert(NW % NG == 0, "warps per group");
    // TM: a window's band cells arrive as one TMA box of S rows x BC columns (the window plus the slack of a
    // 16-byte aligned start) in the window's staging slot; behind the box a scratch area holds the lower-left
    // block for the filters; the float64 window later overwrites both in place.
    static constexpr int BC = (S + 3 + 3) & ~3;
    static constexpr int BOX_BYTES = S * BC * 4;
    static constexpr int SCR_OFF = BOX_BYTES, SCR_BYTES = (W * W + 1) * 8;
    static constexpr int SLOT_raw = F * 8 > SCR_OFF + SCR_BYTES ? F * 8 : SCR_OFF + SCR_BYTES;
    // bytes between staged windows. TM: slots are 128-byte aligned (TMA destination), so the same cell of every
    // window would fall into the same bank; the float64 window and the scratch block are therefore shifted
    // inside the slot by a few words that depend on the window's index (VOFF / SOFF below)
    static constexpr int WSTRIDE = TM ? ((SLOT_raw + 127) & ~127) : F * 8;
    static constexpr int VSLACK = WSTRIDE - F * 8, SSLACK = WSTRIDE - SCR_OFF - SCR_BYTES;
    static constexpr int VSTEP = !TM ? 0 : (VSLACK / 3 >= 40 ? ((VSLACK / 3) & ~7) : (VSLACK & ~7));
    static constexpr int VMASK = VSLACK / 3 >= 40 ? 3 : 1;
    static constexpr int SSTEP = !TM ? 0 : ((SSLACK / 3) & ~7);
    static constexpr int NBUF = TM == 2 ? 2 : 1;              // TM == 2: the boxes of the next take land while this one is processed
    static constexpr int PB_raw = (int)(stage_bytes / NG / NBUF / (size_t)WSTRIDE) & ~1;
    static constexpr int PB = PB_raw < P ? PB_raw : P;        // windows staged per take and group (even)
    static constexpr int NIT = (PB * S + NTG - 1) / NTG;      // TM: window columns per thread and take
    static constexpr int NIW = (PB + 31) / 32;                // TM: warps of a group that request boxes
    static_assert(!TM || NIW * 32 <= NTG, "issuing warps");
    static constexpr size_t fea_bytes = (size_t)P * F * 4;
    static_assert(F * 4 >= S * 16, "row extrema do not fit a feature row");
    static constexpr int NPW = (PB / 2 + NWG - 1) / NWG;      // window pairs per warp per take
    static size_t total(int ND, int n_trees) {
        return stage_bytes + fea_bytes + 2 * (size_t)((ND + 1) & ~1) * 8 + (size_t)NG * NBUF * PB * 4 + 2 * (size_t)F * 8 +
               (size_t)P * 4 + 2 * (size_t)NG * NBUF * PB * 4 + (size_t)n_trees * 4 + (size_t)((n_trees + 3) & ~3) +
               2 * (size_t)NG * PB * 2 + (size_t)P + 64 + (CF ? (size_t)((n_trees + 3) & ~3) : 0) +
               (TM ? (size_t)NG * NBUF * PB * 6 + 16 : 0);
    }
};

// order-preserving map double -> uint64 (integer min / max reductions) and back
__device__ __forceinline__ unsigned long long pk_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double pk_unkey(unsigned long long k) {
    return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull)));
}

// explicit shared-space accesses on 32-bit addresses: [reg + immediate], no generic-pointer arithmetic
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ void lds_node(uint32_t addr, uint2& nd) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(nd.x), "=r"(nd.y) : "r"(addr));
}
// one level of one chain. `nd` is the node at shared address `addr`; an internal node
// moves to a child and loads it, a leaf stays (its 8 bytes are the leaf value; `addr` keeps
// moving by garbage after a leaf, but every later load is predicated off, so it is never used).
// No branch: the loads are predicated on "internal", the step is selected. NaN features
// are handled by the caller on a separate path (NaN compares false: always right).
//   feature byte offset = y & 0xFFC, right-child byte offset = (y >> 16) - 0x8000 (pk_common.cuh)
__device__ __forceinline__ void pk_step(uint32_t xrow_addr, uint32_t& addr, uint2& nd) {
    asm volatile(
        "{\n"
        ".reg .pred q, le;\n"
        ".reg .u32 t, s;\n"
        ".reg .f32 x, thr;\n"
        "setp.lt.s32 q, %2, 0;\n"
        "and.b32 t, %2, 0xFFC;\n"
        "add.u32 t, t, %3;\n"
        "@q ld.shared.f32 x, [t];\n"
        "mov.b32 thr, %1;\n"
        "setp.le.f32 le, x, thr;\n"
        "shr.u32 s, %2, 16;\n"
        "selp.u32 s, 0x8008, s, le;\n"
        "add.u32 s, s, %0;\n"
        "add.u32 %0, s, 0xFFFF8000;\n"
        "@q ld.shared.v2.u32 {%1, %2}, [%0];\n"
        "}"
        : "+r"(addr), "+r"(nd.x), "+r"(nd.y)
        : "r"(xrow_addr));
}

// The same step on the child-feature encoding (pk_common.cuh): the thread holds the node AND the
// value `xv` of the feature that node tests, so the compare needs no load; the node names the
// features of both children, so the child's node and the child's feature value are fetched together:
// one shared-memory round trip per level instead of two.
//   left feature = y & 0xFF, right feature = (y >> 8) & 0xFF, right-child byte offset = (y >> 16) - 0x8000
__device__ __forceinline__ void pk_step_cf(uint32_t xrow_addr, uint32_t& addr, uint2& nd, uint32_t& xv) {
    asm volatile(
        "{\n"
        ".reg .pred q, le;\n"
        ".reg .u32 t, s, sel;\n"
        ".reg .f32 x, thr;\n"
        "setp.lt.s32 q, %2, 0;\n"
        "mov.b32 thr, %1;\n"
        "mov.b32 x, %3;\n"
        "setp.le.f32 le, x, thr;\n"
        "selp.b32 sel, 0x4440, 0x4441, le;\n"
        "prmt.b32 t, %2, 0, sel;\n"
        "mad.lo.u32 t, t, 4, %4;\n"
        "shr.u32 s, %2, 16;\n"
        "selp.u32 s, 0x8008, s, le;\n"
        "add.u32 s, s, %0;\n"
        "add.u32 %0, s, 0xFFFF8000;\n"
        "@q ld.shared.v2.u32 {%1, %2}, [%0];\n"
        "@q ld.shared.u32 %3, [t];\n"
        "}"
        : "+r"(addr), "+r"(nd.x), "+r"(nd.y), "+r"(xv)
        : "r"(xrow_addr));
}

#ifdef PK_FUSED_CLOCK
#define PK_TICK(k) do { if (tid == 0) { const long long t_ = clock64(); clk[k] += t_ - t0; t0 = t_; } } while (0)
#else
#define PK_TICK(k) do { } while (0)
#endif

template <int W, int P, int TPP, int TBN, int CH, int OCC, int NTH, int NGR, int XSTAGE, int CF, int TM>
__global__ void __launch_bounds__(NTH, OCC) k_score_fused(const __grid_constant__ FusedParams prm) {
    using Cfg = FusedCfg<W, P, TPP, TBN, CH, NTH, NGR, XSTAGE, CF, TM>;
    constexpr int S = Cfg::S, F = Cfg::F, NT = Cfg::NT, NW = Cfg::NW, NS = Cfg::NS, NM = Cfg::NM, CHUNK = Cfg::CHUNK;
    constexpr int PB = Cfg::PB, NG = Cfg::NG, NTG = Cfg::NTG, NWG = Cfg::NWG, WSTRIDE = Cfg::WSTRIDE, NBUF = Cfg::NBUF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [tree buffers | hand-over]  (= window staging during phase A) | features | exp | 1/exp |
    //         candidate rank | gather order | slot->candidate | window distance |
    //         window nonzeros | tree roots | tree depths | kept windows of a take | nan flags | barriers
    uint2* s_nodes = reinterpret_cast<uint2*>(smem_raw);
    double* s_hand = reinterpret_cast<double*>(smem_raw + Cfg::node_bytes);
    double* s_V = reinterpret_cast<double*>(smem_raw);            // [NG][PB][F] float64, phase A only
    float* s_fea = reinterpret_cast<float*>(smem_raw + Cfg::stage_bytes);
    double* s_exp = reinterpret_cast<double*>(smem_raw + Cfg::stage_bytes + Cfg::fea_bytes);
    const int ND = prm.ND, NDp = (ND + 1) & ~1;
    double* s_rexp = s_exp + NDp;                                 // RN(1 / exp)
    int32_t* s_rank = reinterpret_cast<int32_t*>(s_rexp + NDp);   // [NG][PB] candidate rank (PB even)
    int2* s_cell = reinterpret_cast<int2*>(s_rank + NG * NBUF * PB);     // [2F] gather order of a window pair
    int32_t* s_idx = reinterpret_cast<int32_t*>(s_cell + 2 * F);  // [P]
    int32_t* s_cd = s_idx + P;                                    // [NG][PB]
    int32_t* s_nz = s_cd + NG * NBUF * PB;                        // [NG][NBUF][PB]
    uint32_t* s_root = reinterpret_cast<uint32_t*>(s_nz + NG * NBUF * PB);   // [n_trees]
    uint8_t* s_depth = reinterpret_cast<uint8_t*>(s_root + prm.n_trees);     // [n_trees] (padded to 4)
    uint16_t* s_kl = reinterpret_cast<uint16_t*>(s_depth + ((prm.n_trees + 3) & ~3));    // [NG][PB] kept windows of a take
    uint16_t* s_ks = s_kl + NG * PB;                              // [NG][PB] their feature slots
    uint8_t* s_nan = reinterpret_cast<uint8_t*>(s_ks + NG * PB);  // [P]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_nan + P) + 15) & ~(uintptr_t)15);
    uint8_t* s_rootfeat = reinterpret_cast<uint8_t*>(s_bar + 4);  // [n_trees] (CF only)
    int32_t* s_cx = reinterpret_cast<int32_t*>(s_rootfeat + (CF ? ((prm.n_trees + 3) & ~3) : 0));   // [NG][PB] window row (TM only)
    int16_t* s_slot = reinterpret_cast<int16_t*>(s_cx + NG * NBUF * PB);  // [NG][NBUF][PB] feature slot of a kept window, else -1 (TM only)
    __shared__ uint64_t s_gbar[NG * NBUF];                        // TM: the boxes of a group's take have landed
    // phase B: leaf hand-over, running sums, list of pixels still walking
    double* s_lv = s_hand;                                        // [2][CH][P] (TPP == 2)
    double* s_acc = s_hand + (TPP == 2 ? 2 * CH * P : 0);         // [P]
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_acc + P);    // [2][P]
    __shared__ int s_gnkt[NG], s_gtake[NG * NBUF], s_gstop[NG * NBUF], s_reserved, s_nkept, s_done, s_expbad, s_wc[32];
    __shared__ long long s_gstart[NG * NBUF];
    constexpr int GCACHE = 64;                   // tree-group table kept in shared memory when it fits
    __shared__ int4 s_grp[GCACHE];

    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int half = lane >> 4, h = lane & 15;
    const int grp = tid / NTG, gtid = tid - grp * NTG, gw = gtid >> 5;      // feature-building group
    auto gsync = [&]() {
        if (NG == 1) __syncthreads();
        else if (NTG == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(NTG) : "memory");
    };
    const int G = prm.n_groups;
    const long long n_cand = min(prm.ncand_dev[0], prm.cand_cap);

    // ---- one-time setup -----------------------------------------------------
    if (tid == 0) s_expbad = prm.balanced ? (prm.flags[2] & 4) : 0;     // bit 2: a weight outside [1e-45, 1e45]
    __syncthreads();
    for (int i = tid; i < ND; i += NT) {
        const double e = prm.expv[i];
        s_exp[i] = e;
        s_rexp[i] = __ddiv_rn(1.0, e);
        if (!pk_div_safe(e)) s_expbad = 1;
    }
    for (int i = tid; i < prm.n_trees; i += NT) { s_root[i] = prm.roots[i]; s_depth[i] = prm.depth[i]; }
    for (int i = tid; i < G && i < GCACHE; i += NT) s_grp[i] = prm.groups[i];
    if (CF)
        for (int i = tid; i < prm.n_trees; i += NT) s_rootfeat[i] = prm.rootfeat[i];
    auto group_of = [&](int gi) -> int4 { return (G <= GCACHE) ? s_grp[gi] : prm.groups[gi]; };
    // Gather order of a window pair: cells by window diagonal (b - a), then along it -- contiguous
    // in the band. Entry idx of [0, 2F): .x = band offset of the cell relative to (d * pitch + x - W),
    // .y = byte offset in the pair's staging area | lane holding the row weight << 16 |
    //      lane holding the column weight << 21 | window << 26.
    for (int idx = tid; idx < 2 * F; idx += NT) {
        const int k = idx >= F, kk = idx - k * F;
        int a = 0, b = 0, run = 0;
        for (int df = -(S - 1); df <= S - 1; ++df) {
            const int len = S - (df < 0 ? -df : df);
            if (kk < run + len) { a = (df < 0 ? -df : 0) + (kk - run); b = a + df; break; }
            run += len;
        }
        s_cell[idx] = make_int2((int)((long long)(b - a) * prm.pitch + a), ((k * F + a * S + b) * 8) | ((k * 16 + a) << 16) | ((k * 16 + b) << 21) | (k << 26));
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        if (TM)
            for (int g = 0; g < NG * NBUF; ++g) mbar_init(&s_gbar[g], Cfg::NIW);
        mbar_fence_init();
        s_done = 0;
    }
    __syncthreads();
    // Tree groups are streamed through the two buffers; positions count group loads since the
    // start of the kernel (uniform across threads). No load is in flight during phase A.
    uint32_t issued = 0, consumed = 0;
    // Issued by the first lane of the last warp: that warp never accumulates (it belongs to the last
    // sub-thread) and is idle once pixels retire, so refills stay off the CTA's critical path.
    constexpr int ISSUER = NT - 32;
    auto issue = [&](uint32_t pos) {
        if (tid == ISSUER) {
            const int4 g = group_of((int)(pos % G));
            const uint32_t bytes = (uint32_t)(g.w < 0 ? -g.w : g.w) * 8u;
            uint64_t* bar = &s_bar[pos & 1];
            mbar_expect_tx(bar, bytes);
            bulk_g2s(s_nodes + (size_t)(pos & 1) * TBN, prm.nodes + g.z, bytes, bar);
        }
    };

    int cur = 0;                                 // TM == 2: which of the group's two slot sets holds the take being processed
    const uint32_t V_base = smem_u32(s_V) + (uint32_t)(grp * NBUF * PB) * (uint32_t)WSTRIDE;   // this group's staging buffer(s)
    uint32_t V_addr = V_base;
    auto win_base = [&](int i) -> uint32_t {     // float64 window of the take's window i
        return V_addr + (uint32_t)i * (uint32_t)WSTRIDE + (uint32_t)((i & Cfg::VMASK) * Cfg::VSTEP);
    };
    auto scr_base = [&](int i) -> uint32_t {     // TM: parked lower-left block + centre of window i
        return V_addr + (uint32_t)i * (uint32_t)WSTRIDE + (uint32_t)(Cfg::SCR_OFF + ((i >> 2) & 3) * Cfg::SSTEP);
    };
    int32_t* g_cx = s_cx + grp * NBUF * PB;
    int16_t* g_slot = s_slot + grp * NBUF * PB;
    uint32_t gphase = 0;                         // TM: parity of the group's box barrier(s), one bit per slot set
    int32_t* g_rank = s_rank + grp * NBUF * PB;
    int32_t* g_cd = s_cd + grp * NBUF * PB;
    int32_t* g_nz = s_nz + grp * NBUF * PB;
    uint16_t* g_kl = s_kl + grp * PB;
    uint16_t* g_ks = s_ks + grp * PB;
    const uint32_t exp_addr = smem_u32(s_exp), rexp_addr = smem_u32(s_rexp), fea_addr = smem_u32(s_fea);

    bool last = false;
#ifdef PK_FUSED_CLOCK
    long long clk[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t0 = clock64();
    long long clkq[6] = {0, 0, 0, 0, 0, 0};      // inside a forest chunk (thread 0): setup | walk | hand-over + mbarrier | CTA barrier | refill | accumulate
    long long clkg[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // forest phase, per tree group (re-pack time lands in the next group)
#endif
    for (;;) {
        // ================= phase A: features =================
        // Windows are staged PB at a time ("a take") in the shared memory the tree buffers use during
        // phase B. Each group of warps runs its own takes; every step below runs over the whole take
        // with all threads of the group.
        if (tid == 0) { s_reserved = 0; s_nkept = 0; }
        __syncthreads();
        // reserve feature slots, then candidates, for the take that slot set `set` will hold (thread 0 of the group)
        auto grab = [&](int set) {
            const int prev = atomicAdd(&s_reserved, PB);
            const int sz = max(0, min(PB, P - prev));
            int take = 0;
            long long st = 0;
            bool stop = true;
            if (sz > 0) {
                st = (long long)atomicAdd(prm.next, (unsigned long long)sz);
                const long long rem = n_cand - st;
                take = rem <= 0 ? 0 : (int)(rem < sz ? rem : sz);
                stop = rem <= sz;                      // every candidate has been handed out
                if (stop) s_done = 1;
            }
            s_gstart[grp * NBUF + set] = st;
            s_gtake[grp * NBUF + set] = take;
            s_gstop[grp * NBUF + set] = stop;
        };
        auto use_set = [&](int set) {                  // the slot set whose take is processed
            cur = set;
            V_addr = V_base + (uint32_t)(set * PB) * (uint32_t)WSTRIDE;
            g_rank = s_rank + (grp * NBUF + set) * PB;
            g_cd = s_cd + (grp * NBUF + set) * PB;
            g_nz = s_nz + (grp * NBUF + set) * PB;
            g_cx = s_cx + (grp * NBUF + set) * PB;
            g_slot = s_slot + (grp * NBUF + set) * PB;
        };
        // TM: one lane per window writes the take's coordinates and requests the window's band cells from the TMA
        // unit: rows x-W..x+W, columns from (y-W) & ~3 of the dense-matrix view of the row-major band (one
        // cp.async.bulk.tensor.2d per window, SASS UTMALDG.2D), completing on the slot set's mbarrier. Called by the
        // whole issuing warps with the window's coordinates already in registers.
        auto request = [&](int set, int take, int i, int x, int d, int rank) {
            const int o = (grp * NBUF + set) * PB;
            const bool ok = i < take && (x - W >= 0) && (x + d + W + 1 <= prm.n);             // scoreUtils.py:75
            if (i < take) {
                s_rank[o + i] = rank;
                s_cd[o + i] = d; s_cx[o + i] = x; s_nz[o + i] = ok ? 0 : -1; s_slot[o + i] = -1;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            // the slots were last touched through the generic proxy (an earlier take, the forest phase)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) mbar_expect_tx(&s_gbar[grp * NBUF + set], (uint32_t)__popc(m) * (uint32_t)Cfg::BOX_BYTES);
            __syncwarp();
            if (ok) tma_box_2d(V_base + (uint32_t)(set * PB + i) * (uint32_t)WSTRIDE, &prm.tmap, (x + d - W) & ~3, x - W,
                               smem_u32(&s_gbar[grp * NBUF + set]));
        };
        if constexpr (TM == 2) {
            // prologue: the first take of the batch is requested here; inside the loop every take requests its successor
            use_set(0);
            if (gtid == 0) grab(0);
            gsync();
            if (gtid < Cfg::NIW * 32) {
                const int take0 = s_gtake[grp * NBUF], i = gtid;
                const long long start0 = s_gstart[grp * NBUF];
                int x = 0, d = 0, rk = 0;
                if (i < take0) { x = prm.cx[start0 + i]; d = prm.cd[start0 + i]; rk = __ldg(prm.crank + start0 + i); }
                request(0, take0, i, x, d, rk);
            }
        }
        for (;;) {
            if constexpr (TM == 2) {
                if (gtid == 0) s_gnkt[grp] = 0;
            } else {
                if (gtid == 0) { grab(0); s_gnkt[grp] = 0; }
            }
            gsync();
            PK_TICK(0);
            const int take = s_gtake[grp * NBUF + cur];
            const long long start = s_gstart[grp * NBUF + cur];
            const bool stop = s_gstop[grp * NBUF + cur] != 0;
            // TM == 2: the successor take is reserved now (the answer is needed two barriers further down) ...
            if (TM == 2 && gtid == 0) {
                if (!stop) grab(cur ^ 1);
            }
            int nx_x = 0, nx_d = 0, nx_rank = 0, nx_take = 0;      // ... its coordinates are loaded after the column step ...
            if constexpr (TM) {
            if constexpr (TM == 1) {
            // ---- T1 (TM == 1): coordinates and box requests of this take
            if (gtid < Cfg::NIW * 32) {
                const int i = gtid;
                int x = 0, d = 0, rk = 0;
                if (i < take) { x = prm.cx[start + i]; d = prm.cd[start + i]; rk = __ldg(prm.crank + start + i); }
                request(0, take, i, x, d, rk);
            }
            gsync();
            }
            PK_TICK(8);
            // ---- A3a (TM): one thread per window column: counts from the box, balanced values into registers,
            //      non-zeros counted, the lower-left block and the centre parked for the filters
            double v[Cfg::NIT][S];
            {
                bool waited = false;
#pragma unroll
                for (int k = 0; k < Cfg::NIT; ++k) {
                    const int item = gtid + k * NTG;
                    const int i = item / S, b = item - i * S;
#pragma unroll
                    for (int a = 0; a < S; ++a) v[k][a] = 0.0;
                    const bool have = item < take * S && g_nz[i] >= 0;          // border windows were not requested
                    int x = 0, d = 0;
                    double wr[S], wcb = 1.0;
#pragma unroll
                    for (int a = 0; a < S; ++a) wr[a] = 1.0;
                    if (have) {
                        x = g_cx[i]; d = g_cd[i];
                        if (prm.balanced) {
                            wcb = __ldg(prm.w + (x + d - W + b));
#pragma unroll
                            for (int a = 0; a < S; ++a) wr[a] = __ldg(prm.w + (x - W + a));
                        }
                    }
                    if (!waited) { PK_TICK(9); mbar_wait(&s_gbar[grp * NBUF + cur], (gphase >> cur) & 1u); waited = true; PK_TICK(10); }
                    if (have) {
                        const int y0 = x + d - W;
                        const uint32_t cell0 = V_addr + (uint32_t)i * (uint32_t)WSTRIDE + (uint32_t)((y0 & 3) + b) * 4u;
                        int nz = 0;
                        // weight products w[r] * w[c] of the column; the largest high word (as unsigned) tells whether
                        // every product is positive, finite and small enough that product * count cannot overflow
                        double p[S];
                        unsigned hmax = 0u;
#pragma unroll
                        for (int a = 0; a < S; ++a) {
                            p[a] = __dmul_rn(wr[a], wcb);
                            hmax = max(hmax, (unsigned)__double2hiint(p[a]));
                        }
                        if (hmax < 0x7DF00000u && d + b >= S - 1) {
                            // common case: (w[r] w[c]) count needs no finiteness test (a zero count gives +0) and every
                            // cell of the column lies on or above the main diagonal
#pragma unroll
                            for (int a = 0; a < S; ++a) {
                                const double q = __dmul_rn(p[a], (double)(int)lds_u32(cell0 + a * Cfg::BC * 4));     // pk_value
                                v[k][a] = q;
                                nz += q != 0.0;
                            }
                        } else {
#pragma unroll
                            for (int a = 0; a < S; ++a) {
                                int cnt = (int)lds_u32(cell0 + a * Cfg::BC * 4);
                                // below the main diagonal (only when d < 2W): the symmetric cell, read directly
                                if (d + b - a < 0) cnt = __ldg(prm.band2 + ((long long)(y0 + b) * prm.P2 + (a - b - d)));
                                double q = __dmul_rn(p[a], (double)cnt);                  // pk_value
                                q = isfinite(q) ? q : 0.0;
                                if (cnt == 0) q = 0.0;
                                v[k][a] = q;
                                nz += q != 0.0;
                            }
                        }
                        atomicAdd(&g_nz[i], nz);
                        const uint32_t scr = scr_base(i);
                        if (b < W) {
#pragma unroll
                            for (int a = 0; a < W; ++a) sts_f64(scr + (uint32_t)(a * W + b) * 8u, v[k][a]);
                        }
                        if (b == W) sts_f64(scr + (uint32_t)(W * W) * 8u, v[k][W]);
                    }
                }
                gphase ^= 1u << cur;
            }
            PK_TICK(11);
            gsync();
            PK_TICK(1);
            if constexpr (TM == 2) {
                // ... (the reservation made at the top of the loop is visible now; the loads fly during the filter step) ...
                if (!stop && gtid < Cfg::NIW * 32) {
                    nx_take = s_gtake[grp * NBUF + (cur ^ 1)];
                    const long long nstart = s_gstart[grp * NBUF + (cur ^ 1)];
                    if (gtid < nx_take) { nx_x = prm.cx[nstart + gtid]; nx_d = prm.cd[nstart + gtid]; nx_rank = __ldg(prm.crank + nstart + gtid); }
                }
            }
            // ---- A2 (TM): the reference's filters, one thread per window
            {
                const int i = lane * NWG + gw;
                if (i < take) {
                    const int nz = g_nz[i];
                    bool ok = nz >= 0 && !((double)nz < (double)F * 0.1);        // utils.py:225
                    const uint32_t scr = scr_base(i);
                    if (ok) {
                        double s = 0.0;                                            // utils.py:228 (numba order)
#pragma unroll
                        for (int q = 0; q < W * W; ++q) s = __dadd_rn(s, lds_f64(scr + (uint32_t)q * 8u));
                        const double ll = __ddiv_rn(s, (double)(W * W));
                        ok = (ll > 0.0) && (__ddiv_rn(lds_f64(scr + (uint32_t)(W * W) * 8u), ll) > 0.1);  // utils.py:229-232
                    }
                    if (ok) {
                        const int kl = atomicAdd(&s_gnkt[grp], 1);
                        const int slot = atomicAdd(&s_nkept, 1);
                        const long long ci = start + i;
                        g_kl[kl] = (uint16_t)i;
                        g_ks[kl] = (uint16_t)slot;
                        g_slot[i] = (int16_t)slot;
                        s_idx[slot] = (int)ci;
                        prm.keep[ci] = 1;
                        atomicAdd(&prm.batch_win[g_rank[i] / PK_BATCH], 1);
                        s_nan[slot] = 0;
                    }
                }
            }
            gsync();
            PK_TICK(2);
            {
                const int nkt_ = s_gnkt[grp];
                if (gtid == 0 && nkt_ < take) atomicSub(&s_reserved, take - nkt_);    // rejected windows free their slots
            }
            if constexpr (TM == 2) {
                // ... and its boxes are requested into the other slot set, whose last reader was the previous take
                if (!stop && gtid < Cfg::NIW * 32) request(cur ^ 1, nx_take, gtid, nx_x, nx_d, nx_rank);
            }
            const bool fastdiv = !s_expbad;
            // ---- A3b (TM): distance normalisation + vertical Gaussian pass of the kept windows, from the registers,
            //      written as the float64 window over the box (every box cell was read before the barriers above)
#pragma unroll
            for (int k = 0; k < Cfg::NIT; ++k) {
                const int item = gtid + k * NTG;
                const int i = item / S, b = item - i * S;
                if (item >= take * S || g_slot[i] < 0) continue;
                const int d = g_cd[i];
                const uint32_t col_addr = win_base(i) + (uint32_t)b * 8u;     // V[a][b] = col_addr + a*S*8
                double u[S], g[S];
                // utils.py:187-200: V[a][b] / exp[|d + b - a|]
                if (fastdiv && d + b >= S - 1) {
                    const uint32_t e0 = exp_addr + (uint32_t)(d + b) * 8u, r0 = rexp_addr + (uint32_t)(d + b) * 8u;
#pragma unroll
                    for (int a = 0; a < S; ++a) u[a] = pk_div_r(v[k][a], lds_f64(e0 - a * 8), lds_f64(r0 - a * 8));
                } else if (fastdiv) {
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        int dd = d + b - a;
                        dd = dd < 0 ? -dd : dd;
                        u[a] = pk_div_r(v[k][a], lds_f64(exp_addr + dd * 8), lds_f64(rexp_addr + dd * 8));
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        int dd = d + b - a;
                        dd = dd < 0 ? -dd : dd;
                        u[a] = __ddiv_rn(v[k][a], lds_f64(exp_addr + dd * 8));
                    }
                }
#pragma unroll
                for (int a = 0; a < S; ++a) {
                    double t = __dmul_rn(u[a], PK_GK[4]);
#pragma unroll
                    for (int jj = 4; jj >= 1; --jj)
                        t = __dadd_rn(t, __dmul_rn(__dadd_rn(u[pk_reflect(a - jj, S)], u[pk_reflect(a + jj, S)]), PK_GK[4 - jj]));
                    g[a] = t;
                }
#pragma unroll
                for (int a = 0; a < S; ++a) sts_f64(col_addr + a * S * 8, g[a]);
            }
            gsync();
            PK_TICK(3);
            } else {
            // ---- A1: gather + balance. A warp owns NPW window pairs of the take; all 32 lanes share the
            //      2*F cells of a pair. Every band load of the take is issued before the first is used.
            {
                constexpr int NPW = Cfg::NPW;
                int myx = 0, myd = 0;
                bool myok = false;
                if (lane < 2 * NPW) {
                    const int i = 2 * (gw + (lane >> 1) * NWG) + (lane & 1);
                    if (i < take) {
                        myx = prm.cx[start + i];
                        myd = prm.cd[start + i];
                        myok = (myx - W >= 0) && (myx + myd + W + 1 <= prm.n);         // scoreUtils.py:75
                        g_rank[i] = __ldg(prm.crank + start + i);
                    }
                }
#ifdef PK_FUSED_CLOCK
                if (__shfl_sync(0xffffffffu, myx, 0) == -12345) clk[12] = 1;
                PK_TICK(8);
#endif
                // this lane's cells of a pair (the same for every pair)
                int cellx[NS], celly[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int2 cell = s_cell[min(s * 32 + lane, 2 * F - 1)];
                    cellx[s] = cell.x; celly[s] = cell.y;
                }
                int cnt[NPW][NS];
                double wr[NPW], wc[NPW];
#pragma unroll
                for (int q = 0; q < NPW; ++q) {
                    const int x0 = __shfl_sync(0xffffffffu, myx, 2 * q), d0 = __shfl_sync(0xffffffffu, myd, 2 * q);
                    const int x1 = __shfl_sync(0xffffffffu, myx, 2 * q + 1), d1 = __shfl_sync(0xffffffffu, myd, 2 * q + 1);
                    const bool ok0 = __shfl_sync(0xffffffffu, (int)myok, 2 * q), ok1 = __shfl_sync(0xffffffffu, (int)myok, 2 * q + 1);
                    // every cell of the window lies on a stored diagonal d + (b - a) in [0, ND - 2]
                    const bool fa0 = d0 >= S - 1 && d0 + S - 1 < ND - 1, fa1 = d1 >= S - 1 && d1 + S - 1 < ND - 1;
                    const int base0 = d0 * (int)prm.pitch + (x0 - W), base1 = d1 * (int)prm.pitch + (x1 - W);
                    if (ok0 && ok1 && fa0 && fa1) {                         // uniform in the warp: the common case
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            // window of this lane's cell: known at compile time except in the slot that straddles F
                            const bool k = (s * 32 + 31 < F) ? false : ((s * 32 >= F) ? true : (lane >= F - s * 32));
                            cnt[q][s] = 0;
                            if (s * 32 + 31 < 2 * F || s * 32 + lane < 2 * F)
                                cnt[q][s] = __ldg(prm.band + ((k ? base1 : base0) + cellx[s]));             // scoreUtils.py:31
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            const int idx = s * 32 + lane;
                            cnt[q][s] = 0;
                            const int k = (celly[s] >> 26) & 1;
                            if (idx < 2 * F && (k ? ok1 : ok0)) {
                                const int a = (celly[s] >> 16) & 15, b = (celly[s] >> 21) & 15;
                                const int xx = k ? x1 : x0, dd0 = k ? d1 : d0;
                                const int r = xx - W + a, c = xx + dd0 - W + b;
                                const int dd = c - r, ad = dd < 0 ? -dd : dd, lo = dd < 0 ? c : r;
                                if (ad < ND - 1) cnt[q][s] = __ldg(prm.band + (long long)ad * prm.pitch + lo);
                            }
                        }
                    }
                    // lane h of a half holds the row weight w[x - W + h] and the column weight w[x + d - W + h]
                    wr[q] = 0.0; wc[q] = 0.0;
                    const bool okh = half ? ok1 : ok0;
                    if (prm.balanced && okh && h < S) {
                        const int xh = half ? x1 : x0, dh = half ? d1 : d0;
                        wr[q] = __ldg(prm.w + (xh - W + h));
                        wc[q] = __ldg(prm.w + (xh + dh - W + h));
                    }
                }
                PK_TICK(9);
#pragma unroll
                for (int q = 0; q < NPW; ++q) {
                    const int j0 = 2 * (gw + q * NWG);
                    if (j0 >= take) break;                                    // uniform in the warp
                    const uint32_t V0_addr = V_addr + (uint32_t)(j0 * F) * 8u;
                    int nzp = 0;                                              // nonzeros: window 0 | window 1 << 16
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const bool valid = s * 32 + 31 < 2 * F || s * 32 + lane < 2 * F;
                        double v = (double)cnt[q][s];
                        if (prm.balanced) {
                            const double wa = __shfl_sync(0xffffffffu, wr[q], (celly[s] >> 16) & 31);
                            const double wb = __shfl_sync(0xffffffffu, wc[q], (celly[s] >> 21) & 31);
                            v = __dmul_rn(__dmul_rn(wa, wb), v);              // pk_value
                            v = isfinite(v) ? v : 0.0;
                        }
                        if (cnt[q][s] == 0) v = 0.0;
                        // windows that fail the border test are dropped in A2; their cells are all zero
                        if (valid) {
                            sts_f64(V0_addr + (uint32_t)(celly[s] & 0xFFFF), v);
                            if (v != 0.0) nzp += 1 << ((celly[s] >> 22) & 16);
                        }
                    }
                    nzp = __reduce_add_sync(0xffffffffu, nzp);
                    const bool okh = __shfl_sync(0xffffffffu, (int)myok, 2 * q + half) != 0;
                    const int dh = __shfl_sync(0xffffffffu, myd, 2 * q + half);
                    if (h == 0 && j0 + half < take) {
                        g_nz[j0 + half] = okh ? ((nzp >> (16 * half)) & 0xFFFF) : -1;
                        g_cd[j0 + half] = dh;
                    }
                }
                PK_TICK(10);
            }
            gsync();
            PK_TICK(1);
            // ---- A2: the reference's filters, one thread per window (spread over all warps)
            {
                const int i = lane * NWG + gw;
                if (i < take) {
                    const int nz = g_nz[i];
                    bool ok = nz >= 0 && !((double)nz < (double)F * 0.1);        // utils.py:225
                    const uint32_t myV = V_addr + (uint32_t)(i * F) * 8u;
                    if (ok) {
                        double s = 0.0;                                            // utils.py:228 (numba order)
#pragma unroll
                        for (int a = 0; a < W; ++a)
#pragma unroll
                            for (int b = 0; b < W; ++b) s = __dadd_rn(s, lds_f64(myV + (uint32_t)(a * S + b) * 8u));
                        const double ll = __ddiv_rn(s, (double)(W * W));
                        ok = (ll > 0.0) && (__ddiv_rn(lds_f64(myV + (uint32_t)(W * S + W) * 8u), ll) > 0.1);  // utils.py:229-232
                    }
                    if (ok) {
                        const int kl = atomicAdd(&s_gnkt[grp], 1);
                        const int slot = atomicAdd(&s_nkept, 1);
                        const long long ci = start + i;
                        g_kl[kl] = (uint16_t)i;
                        g_ks[kl] = (uint16_t)slot;
                        s_idx[slot] = (int)ci;
                        prm.keep[ci] = 1;
                        atomicAdd(&prm.batch_win[g_rank[i] / PK_BATCH], 1);
                        s_nan[slot] = 0;
                    }
                }
            }
            gsync();
            PK_TICK(2);
            const int nkt = s_gnkt[grp];
            if (gtid == 0 && nkt < take) atomicSub(&s_reserved, take - nkt);    // rejected windows free their slots
            const bool fastdiv = !s_expbad;      // weights and expected values inside pk_div_r's proven range
            // ---- A3: distance normalisation + vertical Gaussian pass, one thread per window column, in place
            for (int item = gtid; item < nkt * S; item += NTG) {
                const int kl = item / S, b = item - kl * S;
                const int i = g_kl[kl];
                const int d = g_cd[i];
                const uint32_t col_addr = V_addr + (uint32_t)(i * F + b) * 8u;     // V[a][b] = col_addr + a*S*8
                double v[S], g[S];
                // utils.py:187-200: V[a][b] / exp[|d + b - a|]
                if (fastdiv && d + b >= S - 1) {
                    // common case: d + b - a >= 0 for every a, so exp is read at fixed offsets
                    const uint32_t e0 = exp_addr + (uint32_t)(d + b) * 8u, r0 = rexp_addr + (uint32_t)(d + b) * 8u;
#pragma unroll
                    for (int a = 0; a < S; ++a)
                        v[a] = pk_div_r(lds_f64(col_addr + a * S * 8), lds_f64(e0 - a * 8), lds_f64(r0 - a * 8));
                } else if (fastdiv) {
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        int dd = d + b - a;
                        dd = dd < 0 ? -dd : dd;
                        v[a] = pk_div_r(lds_f64(col_addr + a * S * 8), lds_f64(exp_addr + dd * 8), lds_f64(rexp_addr + dd * 8));
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        int dd = d + b - a;
                        dd = dd < 0 ? -dd : dd;
                        v[a] = __ddiv_rn(lds_f64(col_addr + a * S * 8), lds_f64(exp_addr + dd * 8));
                    }
                }
#pragma unroll
                for (int a = 0; a < S; ++a) {
                    double t = __dmul_rn(v[a], PK_GK[4]);
#pragma unroll
                    for (int jj = 4; jj >= 1; --jj)
                        t = __dadd_rn(t, __dmul_rn(__dadd_rn(v[pk_reflect(a - jj, S)], v[pk_reflect(a + jj, S)]), PK_GK[4 - jj]));
                    g[a] = t;
                }
#pragma unroll
                for (int a = 0; a < S; ++a) sts_f64(col_addr + a * S * 8, g[a]);
            }
            gsync();
            PK_TICK(3);
            }
            const int nkt = s_gnkt[grp];
            // ---- A4: horizontal pass, one thread per window row, in place
            for (int item = gtid; item < nkt * S; item += NTG) {
                const int kl = item / S, a = item - kl * S;
                const int i = g_kl[kl];
                const uint32_t row_addr = win_base(i) + (uint32_t)(a * S) * 8u;
                double t[S];
#pragma unroll
                for (int b = 0; b < S; ++b) t[b] = lds_f64(row_addr + b * 8);
                double mn = CUDART_INF, mx = -CUDART_INF;
                bool has_nan = false;
#pragma unroll
                for (int b = 0; b < S; ++b) {
                    double q = __dmul_rn(t[b], PK_GK[4]);
#pragma unroll
                    for (int jj = 4; jj >= 1; --jj)
                        q = __dadd_rn(q, __dmul_rn(__dadd_rn(t[pk_reflect(b - jj, S)], t[pk_reflect(b + jj, S)]), PK_GK[4 - jj]));
                    sts_f64(row_addr + b * 8, q);
                    has_nan |= (q != q);
                    mn = (q < mn) ? q : mn;          // a NaN never wins a comparison; has_nan carries it
                    mx = (q > mx) ? q : mx;
                }
                // row extrema as order-preserving keys, parked in the (still unused) feature row of the
                // window's slot; all ones in the max: NaN seen (numba min/max propagate NaN)
                const unsigned long long kmn = pk_key(mn), kmx = has_nan ? ~0ull : pk_key(mx);
                const uint32_t tmp = fea_addr + (uint32_t)(g_ks[kl] * F) * 4u + (uint32_t)a * 16u;
                sts_u32(tmp, (uint32_t)kmn);
                sts_u32(tmp + 4, (uint32_t)(kmn >> 32));
                sts_u32(tmp + 8, (uint32_t)kmx);
                sts_u32(tmp + 12, (uint32_t)(kmx >> 32));
            }
            gsync();
            PK_TICK(4);
            // ---- A5: min-max scaling to float32 features (utils.py:202-207), one warp per window, two
            //      windows interleaved
            for (int kl0 = gw; kl0 < nkt; kl0 += 2 * NWG) {
                double mn[2], range[2], rr[2];
                bool fr[2];
                uint32_t win[2], frow[2];
                int slot[2];
                bool have[2];
                uint32_t klo[2][2], khi[2][2];          // [window][min, max]
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int kl = kl0 + u * NWG;
                    have[u] = kl < nkt;
                    const int klc = have[u] ? kl : kl0;
                    win[u] = win_base(g_kl[klc]) + (uint32_t)lane * 8u;
                    slot[u] = g_ks[klc];
                    frow[u] = fea_addr + (uint32_t)(slot[u] * F) * 4u;
                    klo[u][0] = 0xffffffffu; khi[u][0] = 0xffffffffu; klo[u][1] = 0u; khi[u][1] = 0u;
                    if (lane < S) {
                        const uint32_t tmp = frow[u] + (uint32_t)lane * 16u;
                        klo[u][0] = lds_u32(tmp); khi[u][0] = lds_u32(tmp + 4);
                        klo[u][1] = lds_u32(tmp + 8); khi[u][1] = lds_u32(tmp + 12);
                    }
                    frow[u] += (uint32_t)lane * 4u;
                }
                __syncwarp();                    // the parked extrema have been read; the rows are overwritten below
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    // 64-bit min / max over the warp: high words first, then low words among the ties
                    const uint32_t hmin = __reduce_min_sync(0xffffffffu, khi[u][0]);
                    const uint32_t lmin = __reduce_min_sync(0xffffffffu, khi[u][0] == hmin ? klo[u][0] : 0xffffffffu);
                    const uint32_t hmax = __reduce_max_sync(0xffffffffu, khi[u][1]);
                    const uint32_t lmax = __reduce_max_sync(0xffffffffu, khi[u][1] == hmax ? klo[u][1] : 0u);
                    const unsigned long long kmin = ((unsigned long long)hmin << 32) | lmin, kmax = ((unsigned long long)hmax << 32) | lmax;
                    double mx = pk_unkey(kmax);
                    mn[u] = pk_unkey(kmin);
                    if (kmax == ~0ull) { mn[u] = CUDART_NAN; mx = CUDART_NAN; }
                    range[u] = __dsub_rn(mx, mn[u]);
                    fr[u] = pk_div_safe(range[u]) && mx <= 1e100;              // else: plain IEEE division
                    rr[u] = __ddiv_rn(1.0, range[u]);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (!have[u]) continue;
                    double gv[NM];
#pragma unroll
                    for (int m = 0; m < NM; ++m) gv[m] = (m * 32 + lane < F) ? lds_f64(win[u] + m * 256) : 0.0;
                    if (fr[u]) {
#pragma unroll
                        for (int m = 0; m < NM; ++m)
                            if (m * 32 + lane < F)
                                sts_f32(frow[u] + m * 128, __double2float_rn(pk_div_r(__dsub_rn(gv[m], mn[u]), range[u], rr[u])));
                    } else {
                        bool fnan = false;
#pragma unroll
                        for (int m = 0; m < NM; ++m)
                            if (m * 32 + lane < F) {
                                const double q = __ddiv_rn(__dsub_rn(gv[m], mn[u]), range[u]);
                                fnan |= isnan(q);
                                sts_f32(frow[u] + m * 128, __double2float_rn(q));
                            }
                        // any NaN feature sends this pixel through the missing_go_to_left-aware walk
                        if (fnan) s_nan[slot[u]] = 1;
                    }
                }
            }
            if (stop) break;                     // uniform in the group
            if constexpr (TM == 2) use_set(cur ^ 1);      // the barrier at the top of the loop separates the takes
            else gsync();                        // the staging buffer and the grab variables are free again
            PK_TICK(5);
        }
        __syncthreads();
        const int nkept = s_nkept;
        last = s_done != 0;
        if (prm.fea_tap != nullptr) {
            // parity tap: the float32 feature rows the forest is about to read, spilled per candidate
            for (int i = tid; i < nkept * F; i += NT) {
                const int sl = i / F;
                prm.fea_tap[(size_t)s_idx[sl] * F + (i - sl * F)] = s_fea[i];
            }
        }

        // ================= phase B: forest =================
        if (nkept > 0) {
            // A pixel whose running sum can no longer exceed min_prob (every leaf value is <= 1) stops
            // walking trees: only pixels with prob > min_prob are ever emitted, and theirs stay exact.
            // The survivors are re-packed onto the low threads after each tree group, so whole warps retire.
            const int slot = tid % P, sub = tid / P;       // sub-thread `sub` walks trees [CH*sub, CH*sub+CH) of a chunk
            const double T = (double)prm.n_trees;
            const double die_below = prm.thre * T - 1e-9 * T;              // margin >> rounding of the sums
            const bool prune = prm.thre > 0.0;
            // the staging area is dead: start streaming the forest into it (two groups ahead)
            if (tid == ISSUER) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            {
                const uint32_t target = consumed + (uint32_t)(G < 2 ? G : 2);
                while (issued < target) { issue(issued); ++issued; }
            }
            for (int i = tid; i < nkept; i += NT) { s_list[i] = (uint16_t)i; s_acc[i] = 0.0; }
            int na = nkept, cur = 0, last_pack = 0;
            int lvpar = 0;
            // one warp polls the mbarrier, the others wait at the CTA barrier (no spinning warps)
            if (wib == 0) mbar_wait(&s_bar[consumed & 1], (consumed >> 1) & 1);
            __syncthreads();
#ifdef PK_FUSED_CLOCK
            long long tg0 = clock64();
#endif
            for (int gi = 0; gi < G; ++gi) {
                const uint32_t pos = consumed;
                const int4 grp = group_of(gi);
                const uint32_t gbase = (uint32_t)grp.z;
                const bool fits = grp.w > 0;                   // every tree of the group is fully staged
                const uint32_t staged = (uint32_t)(fits ? grp.w : -grp.w);
                const uint2* buf = s_nodes + (size_t)(pos & 1) * TBN;
                const uint32_t buf_addr = smem_u32(buf);
                const int t_end = grp.x + grp.y;
                for (int tc = grp.x; tc < t_end; tc += CHUNK) {
#ifdef PK_FUSED_CLOCK
                    long long tq0 = clock64();
#define PK_TICKQ(k) do { if (tid == 0) { const long long t_ = clock64(); clkq[k] += t_ - tq0; tq0 = t_; } } while (0)
#else
#define PK_TICKQ(k) do { } while (0)
#endif
                    const int t = tc + CH * sub;               // this thread's CH trees
                    const bool mine = tid < P * TPP && slot < na;
                    const int pix = mine ? s_list[cur * P + slot] : 0;
                    const float* xrow = s_fea + (size_t)pix * F;
                    const uint32_t xrow_addr = fea_addr + (uint32_t)(pix * F) * 4u;
                    const bool warp_nan = __any_sync(0xffffffffu, mine && s_nan[pix]);
                    double lv[CH];
                    bool ex[CH];
#pragma unroll
                    for (int k = 0; k < CH; ++k) { lv[k] = 0.0; ex[k] = t + k < t_end; }
                    if (mine && ex[0]) {
                        if (fits && !warp_nan) {
                            // chains past the end of the group re-walk tree t and are dropped
                            uint32_t addr[CH];
                            uint2 nd[CH];
                            uint32_t xv[CH];                 // CF: value of the feature the current node tests
                            int maxd = 0;
#pragma unroll
                            for (int k = 0; k < CH; ++k) {
                                const int tk = ex[k] ? t + k : t;
                                addr[k] = buf_addr + (s_root[tk] - gbase) * 8u;
                                maxd = max(maxd, (int)s_depth[tk]);
                                xv[k] = CF ? xrow_addr + 4u * s_rootfeat[tk] : 0u;
                            }
#pragma unroll
                            for (int k = 0; k < CH; ++k) {
                                lds_node(addr[k], nd[k]);
                                if (CF) xv[k] = lds_u32(xv[k]);
                            }
                            PK_TICKQ(0);
                            for (int lvl = 0; lvl < maxd; ++lvl) {
#pragma unroll
                                for (int k = 0; k < CH; ++k) {
                                    if (CF) pk_step_cf(xrow_addr, addr[k], nd[k], xv[k]);
                                    else pk_step(xrow_addr, addr[k], nd[k]);
                                }
                            }
#pragma unroll
                            for (int k = 0; k < CH; ++k) lv[k] = __hiloint2double((int)nd[k].y, (int)nd[k].x);
                        } else {
                            // general walk: NaN features follow missing_go_to_left; the tail of a tree
                            // larger than the staging buffer is read from global memory (L2)
#pragma unroll
                            for (int k = 0; k < CH; ++k) {
                                if (!ex[k]) continue;
                                uint32_t p = s_root[t + k] - gbase;
                                uint2 nd = p < staged ? buf[p] : __ldg(prm.nodes + gbase + p);
                                uint32_t ft = CF ? s_rootfeat[t + k] : 0u;        // CF: feature of the current node
                                while ((int)nd.y < 0) {                            // fused-kernel encodings (pk_common.cuh)
                                    const float xv = xrow[CF ? ft : ((nd.y & 0xFFCu) >> 2)];
                                    bool left = xv <= __uint_as_float(nd.x);
                                    if (isnan(xv)) left = CF ? (PK_NODE_MGL(__ldg(&prm.nodes_classic[gbase + p].y)) != 0u) : ((nd.y >> 15) & 1u) != 0u;
                                    if (CF) ft = left ? (nd.y & 0xFFu) : ((nd.y >> 8) & 0xFFu);
                                    p += left ? 1u : ((nd.y >> 19) & 4095u);
                                    nd = p < staged ? buf[p] : __ldg(prm.nodes + gbase + p);
                                }
                                lv[k] = __hiloint2double((int)nd.y, (int)nd.x);
                            }
                        }
                    }
                    // After the last chunk of a group the buffer is handed back: one barrier covers
                    // "everyone is done with it", "the next group has landed" and the leaf hand-over.
                    PK_TICKQ(1);
                    const bool rotate = tc + CHUNK >= t_end;
                    double* lvb = s_lv + (size_t)lvpar * CH * P;
                    if (TPP == 2 && sub == 1 && mine) {
#pragma unroll
                        for (int k = 0; k < CH; ++k) lvb[k * P + slot] = lv[k];
                    }
                    if (rotate && wib == 0 && consumed + 1 < issued) mbar_wait(&s_bar[(consumed + 1) & 1], ((consumed + 1) >> 1) & 1);
                    PK_TICKQ(2);
                    if (TPP == 2 || rotate) __syncthreads();
                    PK_TICKQ(3);
                    if (rotate) {
                        ++consumed;
                        // refill the freed buffer with the group two positions ahead -- unless that group
                        // belongs to the next batch, whose windows are staged here first
                        if (gi + 2 < G) { issue(issued); ++issued; }
                    }
                    PK_TICKQ(4);
                    // ordered accumulation: trees tc .. tc+CHUNK-1 in estimator order
                    if (mine && sub == 0) {
                        double acc = s_acc[pix];
#pragma unroll
                        for (int k = 0; k < CH; ++k)
                            if (ex[k]) acc = __dadd_rn(acc, lv[k]);
                        if (TPP == 2) {
#pragma unroll
                            for (int k = 0; k < CH; ++k)
                                if (tc + CH + k < t_end) acc = __dadd_rn(acc, lvb[k * P + slot]);
                        }
                        s_acc[pix] = acc;
                    }
                    lvpar ^= 1;
                    PK_TICKQ(5);
                }
#ifdef PK_FUSED_CLOCK
                if (tid == 0 && gi < 16) { const long long t_ = clock64(); clkg[gi] += t_ - tg0; tg0 = t_; }
#endif
                // re-pack the pixels that can still exceed min_prob (possible once done > (1 - thre) T)
                if (prune && gi + 1 < G && (double)t_end > T - prm.thre * T && t_end - last_pack >= 12) {
                    last_pack = t_end;
                    __syncthreads();                               // sums of this group are in s_acc
                    const double remaining = T - (double)t_end;
                    bool alive = false;
                    int px = 0;
                    if (tid < na) {
                        px = s_list[cur * P + tid];
                        alive = !(s_acc[px] + remaining < die_below);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, alive);
                    if (lane == 0) s_wc[wib] = __popc(bal);
                    __syncthreads();
                    int pre = 0, tot = 0;
                    for (int k = 0; k < NW; ++k) { const int c = s_wc[k]; if (k < wib) pre += c; tot += c; }
                    if (alive) s_list[(cur ^ 1) * P + pre + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)px;
                    __syncthreads();
                    na = tot;
                    cur ^= 1;
                }
            }
            __syncthreads();
            if (tid < nkept) prm.prob[s_idx[tid]] = __ddiv_rn(s_acc[tid], T);      // partial (< min_prob) for retired pixels
            if (tid == 0) atomicAdd(&prm.counters[1], (unsigned long long)nkept);
        }
        PK_TICK(6);
        if (last) break;
        __syncthreads();
    }
#ifdef PK_FUSED_CLOCK
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 77)) {
        printf("cta %d A1 detail: coords %lld issue %lld process %lld (TM: request %lld weights %lld wait %lld columns %lld)\n", blockIdx.x, clk[8], clk[9], clk[10], clk[8], clk[9], clk[10], clk[11]);
        printf("cta %d forest cycles per tree group: %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", blockIdx.x, clkg[0], clkg[1],
               clkg[2], clkg[3], clkg[4], clkg[5], clkg[6], clkg[7], clkg[8], clkg[9], clkg[10], clkg[11], clkg[12], clkg[13], clkg[14], clkg[15]);
        printf("cta %d forest chunk (thread 0): setup %lld walk %lld handover+mbar %lld barrier %lld refill %lld accumulate %lld\n", blockIdx.x, clkq[0], clkq[1], clkq[2], clkq[3], clkq[4], clkq[5]);
        printf("cta %d cycles: grab %lld A1 %lld A2 %lld A3 %lld A4 %lld A5a %lld A5b %lld B %lld\n", blockIdx.x, clk[0], clk[1], clk[2],
               clk[3], clk[4], clk[7], clk[5], clk[6]);
    }
#endif
    // every issued group has been waited for (issued == consumed after a batch): nothing to drain
}

template <int W, int P, int TPP, int TBN, int CH, int OCC, int NTH = P * TPP, int NGR = 2, int XSTAGE = 0, int CF = 0, int TM = 0>
static int launch_fused_t(const FusedParams& prm_in, pk_forest* f, int ND, int sm_count, cudaStream_t stream) {
    using Cfg = FusedCfg<W, P, TPP, TBN, CH, NTH, NGR, XSTAGE, CF, TM>;
    FusedParams prm = prm_in;
    prm.nodes = CF ? f->d_nodes_f1 : f->d_nodes_f0;
    prm.rootfeat = f->d_rootfeat;
    prm.nodes_classic = f->d_nodes;
    PK_CHECK(pk_forest_groups(f, TBN, Cfg::CHUNK, &prm.groups, &prm.n_groups));
    const size_t smem = Cfg::total(ND, prm.n_trees);
    if (OCC * (smem + 1024) > 228 * 1024) { pk_set_error("fused kernel: %zu bytes of shared memory needed (x%d per SM)", smem, OCC); return PK_EUNSUPPORTED; }
    PK_OPT_IN_SMEM((k_score_fused<W, P, TPP, TBN, CH, OCC, NTH, NGR, XSTAGE, CF, TM>), smem, f->device);
    unsigned grid = (unsigned)(sm_count * OCC);        // persistent: CTAs without work exit at once
    k_score_fused<W, P, TPP, TBN, CH, OCC, NTH, NGR, XSTAGE, CF, TM><<<grid, NTH, smem, stream>>>(prm);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int pk_launch_fused(pk_chrom* c, pk_forest* f, int variant, double thre, int reserve_sms, int child_features, int tma, float* fea_tap) {
    FusedParams prm;
    // windows as TMA boxes of the row-major band copy (default variants only)
    const bool tm = tma && (variant == 0 || variant >= 5) && c->d_band2 && c->band2_valid && c->tmap_ok;
    if (variant >= 5 && variant != 11 && !tm) { pk_set_error("fused variant %d needs the row-major band copy (tuning 'tma')", variant); return PK_EUNSUPPORTED; }
    prm.band2 = c->d_band2; prm.P2 = c->P2;
    if (tm) prm.tmap = c->tmap; else memset(&prm.tmap, 0, sizeof prm.tmap);
    prm.fea_tap = fea_tap;
    prm.band = c->d_band; prm.w = c->d_w; prm.expv = c->d_exp;
    prm.n = c->n; prm.pitch = c->pitch; prm.balanced = c->balanced; prm.ND = c->ND;
    prm.cx = c->d_cx; prm.cd = c->d_cd; prm.crank = c->d_crank;
    prm.ncand_dev = c->d_ncand; prm.cand_cap = c->cand_cap;
    prm.nodes = f->d_nodes; prm.roots = f->d_root; prm.depth = f->d_depth; prm.groups = nullptr;
    prm.n_groups = 0; prm.n_trees = f->n_trees; prm.rootfeat = nullptr;
    // Child-feature encoding (default variants only): one shared-memory round trip per level instead of two,
    // one instruction more per level. Measured (profiles/r1_summary.md): -6 % on the w = 7 kernel (two chains
    // per thread, latency-bound), +3 % on the w = 5 kernel (four chains per thread, closer to issue-bound).
    const bool cf = f->cf_ok && (child_features < 0 ? c->w == 7 : child_features != 0);
    prm.keep = c->d_keep; prm.prob = c->d_prob; prm.batch_win = c->d_batch_win; prm.counters = c->d_counters;
    prm.next = c->d_counters + 2;
    prm.flags = c->d_flags;
    prm.thre = thre;
    if (!f->fused_ok) {         // a right child 4096 or more nodes away: the fused kernel's node encodings cannot hold it
        pk_set_error("fused kernel: a tree of this forest is too large for its node encoding");
        return PK_EUNSUPPORTED;
    }
    if ((long long)c->ND * c->pitch >= (1LL << 31)) {      // the gather indexes the band with 32-bit offsets
        pk_set_error("fused kernel: band of %d x %lld cells exceeds 2^31", c->ND, (long long)c->pitch);
        return PK_EUNSUPPORTED;
    }
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, c->device);
    sm = std::max(1, sm - reserve_sms);
    // variant 0 is the default; the others are tuning experiments (pk_set_tuning("fused", 1 + variant)). Arrangements that
    // were measured and dropped from the build are listed with their times in profiles/r2_summary.md.
    cudaStream_t st = c->stream;
    if (c->w == 5) {
        switch (variant) {
        case 1: return launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 2>(prm, f, c->ND, sm, st);
        case 2: return launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 1>(prm, f, c->ND, sm, st);
        case 3: return launch_fused_t<5, 128, 2, 2112, 2, 2, 256, 2>(prm, f, c->ND, sm, st);      // two CTAs per SM: their phases drift apart and overlap
        case 7: return launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 16, 0, 0, 1>(prm, f, c->ND, sm, st);     // TMA windows, a warp per group
        case 8: return launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 4, 0, 0, 2>(prm, f, c->ND, sm, st);      // TMA windows requested a take ahead (two slot sets), 4 groups
        case 9: return launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 2, 0, 0, 2>(prm, f, c->ND, sm, st);      // same, 2 groups
        default:
            if (tm) return cf ? launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 4, 0, 1, 1>(prm, f, c->ND, sm, st)
                              : launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 4, 0, 0, 1>(prm, f, c->ND, sm, st);
            return cf ? launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 4, 0, 1>(prm, f, c->ND, sm, st)
                      : launch_fused_t<5, 256, 2, 4224, 4, 1, 512, 4>(prm, f, c->ND, sm, st);
        }
    }
    if (c->w == 7) {
        switch (variant) {
        case 1: return launch_fused_t<7, 128, 2, 3200, 2, 1, 384, 2, 32768>(prm, f, c->ND, sm, st);
        case 2: return launch_fused_t<7, 128, 2, 3200, 2, 1, 512, 2, 32768>(prm, f, c->ND, sm, st);
        // eight trees per forest round (four chains per thread, buffers of 6400 nodes) on batches of 112 pixels
        case 8: return launch_fused_t<7, 128, 2, 3200, 2, 1, 384, 1, 32768, 1, 1>(prm, f, c->ND, sm, st);  // TMA windows, four trees per round
        case 14: return launch_fused_t<7, 112, 2, 6400, 4, 1, 384, 3, 0, 1, 2>(prm, f, c->ND, sm, st);     // the default's arrangement with 3 groups of 4 warps
        case 16: return launch_fused_t<7, 112, 2, 6400, 4, 1, 384, 6, 0, 1, 2>(prm, f, c->ND, sm, st);     // ... with 6 groups of 2 warps
        case 17: return launch_fused_t<7, 112, 2, 6400, 4, 1, 384, 1, 0, 1, 1>(prm, f, c->ND, sm, st);     // one group, boxes requested by the take itself
        case 11: return launch_fused_t<7, 112, 2, 6400, 4, 1, 384, 1, 0, 1, 0>(prm, f, c->ND, sm, st);     // eight trees per round on the per-cell gather
        default:
            // TMA windows; with the child-feature encoding also eight trees per forest round (four chains per thread,
            // buffers of 6400 nodes, batches of 112 pixels) and every warp its own group whose boxes are requested a
            // take ahead: 5.53 -> 5.13 -> 4.81 -> 4.59 ms on the C4 chromosome
            if (tm) return cf ? launch_fused_t<7, 112, 2, 6400, 4, 1, 384, 12, 0, 1, 2>(prm, f, c->ND, sm, st)
                              : launch_fused_t<7, 128, 2, 3200, 2, 1, 384, 1, 32768, 0, 1>(prm, f, c->ND, sm, st);
            return cf ? launch_fused_t<7, 128, 2, 3200, 2, 1, 384, 1, 32768, 1>(prm, f, c->ND, sm, st)
                      : launch_fused_t<7, 128, 2, 3200, 2, 1, 384, 1, 32768>(prm, f, c->ND, sm, st);
        }
    }
    return PK_EUNSUPPORTED;
}

bool pk_fused_supported(int w, int n_trees) { return (w == 5 || w == 7) && n_trees <= 2048; }

// ---------------------------------------------------------------------------
// self-test of pk_div_r against IEEE division (pk_selftest_divide)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pk_mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// random double with exponent in [-emax, emax] and a random / adversarial significand
__device__ __forceinline__ double pk_rand_double(unsigned long long r, unsigned long long r2, int emax) {
    unsigned long long mant = r & 0xFFFFFFFFFFFFFull;
    const int mode = (int)(r2 & 7);
    if (mode == 0) mant = 0;                                   // powers of two
    else if (mode == 1) mant = 0xFFFFFFFFFFFFFull;             // all ones
    else if (mode == 2) mant &= 0xFFFFF00000000ull;            // short significands
    else if (mode == 3) mant |= 0xFFFFFFFFull;                 // long runs of ones
    const int e = (int)((r2 >> 8) % (unsigned)(2 * emax + 1)) - emax;
    return __longlong_as_double((long long)(((unsigned long long)(1023 + e) << 52) | mant));
}
__global__ void k_selftest_divide(long long n, unsigned long long seed, unsigned long long* mismatches) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long r0 = pk_mix64(seed + 4 * i), r1 = pk_mix64(seed + 4 * i + 1);
        const unsigned long long r2 = pk_mix64(seed + 4 * i + 2), r3 = pk_mix64(seed + 4 * i + 3);
        const double b = pk_rand_double(r0, r1, 100);
        double a = pk_rand_double(r2, r3, 100);
        const int kind = (int)((r3 >> 40) & 7);
        if (kind == 0) a = b;                                                   // quotient 1
        else if (kind == 1) a = __dmul_rn(b, (double)(1 + (r3 >> 48) % 1000));  // near-exact quotients
        else if (kind == 2) a = __dmul_rn(b, pk_rand_double(r2, r3, 0) * 0.5);  // a in (b/2, b): the min-max case
        else if (kind == 3) a = 0.0;
        if (a != 0.0 && !(a >= 1e-200 && a <= 1e200)) continue;
        const double q = pk_div_r(a, b, __ddiv_rn(1.0, b));
        if (__double_as_longlong(q) != __double_as_longlong(__ddiv_rn(a, b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

int pk_run_selftest_divide(long long n, unsigned long long seed, long long* mismatches) {
    unsigned long long* d = nullptr;
    PK_CUDA(cudaMalloc((void**)&d, 8));
    PK_CUDA(cudaMemset(d, 0, 8));
    k_selftest_divide<<<148 * 8, 256>>>(n, seed, d);
    PK_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    PK_CUDA(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    *mismatches = (long long)h;
    return PK_OK;
}
