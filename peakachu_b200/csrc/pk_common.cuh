// Shared declarations of the peakachu_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "peakachu_b200.h"

#define PK_BATCH 100000            // scoreUtils.py:104
#define PK_MAX_W 12                // window half-width limit: (2w+1)^2 <= 625 features
#define PK_FEAT_BITS 10            // packed node: feature index < 1024

void pk_set_error(const char* fmt, ...);

#define PK_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            pk_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return PK_ECUDA;                                                               \
        }                                                                                  \
    } while (0)

// Opt a kernel into `smem` bytes of dynamic shared memory. The attribute belongs to the (kernel,
// device) pair, so the high-water mark is kept per device: a process may hold handles on several GPUs.
#define PK_MAX_DEVICES 64
#define PK_OPT_IN_SMEM(kernel, smem, device)                                               \
    do {                                                                                   \
        static size_t hw__[PK_MAX_DEVICES] = {0};                                          \
        const int d__ = ((device) >= 0 && (device) < PK_MAX_DEVICES) ? (device) : 0;       \
        if ((size_t)(smem) > 48 * 1024 && (size_t)(smem) > hw__[d__]) {                    \
            PK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem))); \
            hw__[d__] = (size_t)(smem);                                                    \
        }                                                                                  \
    } while (0)

#define PK_CHECK(call)             \
    do {                           \
        int r__ = (call);          \
        if (r__ != PK_OK) return r__; \
    } while (0)

// ---------------------------------------------------------------------------
// forest: nodes renumbered in preorder (left child = parent + 1), 8 bytes each.
//   internal: .x = float32 threshold rounded toward -inf (x_f32 <= t_f64 <=> x_f32 <= .x)
//             .y = 1<<31 | missing_go_left<<30 | (right - self)<<12 [18 bits] | feature*4 [12 bits]
//   leaf:     the 8 bytes are the float64 class-1 fraction (>= 0, so bit 31 of .y is 0)
// ---------------------------------------------------------------------------
struct pk_forest {
    int device = 0;
    int32_t n_trees = 0, n_features = 0;
    int64_t n_nodes = 0;
    uint2* d_nodes = nullptr;        // [n_nodes]
    uint32_t* d_root = nullptr;      // [n_trees] packed index of the root
    int32_t* d_orig = nullptr;       // [n_nodes] tree-local sklearn node id (apply tap)
    uint8_t* d_depth = nullptr;      // [n_trees] depth of the deepest leaf
    // Node arrays of the fused kernel (same numbering, leaves and .x as d_nodes). The right-child byte
    // offset sits in bits 16..30 under the "internal" bit, so `y >> 16` is that offset + 0x8000 and the
    // walk's address update is one three-input add (addr + step - 0x8000). Needs every right offset
    // < 4096 nodes (fused_ok); other forests take the separate feature / forest kernels.
    //   own-feature:    .y = 1<<31 | (right - self)*8 << 16 | missing_go_left << 15 | feature*4
    //   child-feature:  .y = 1<<31 | (right - self)*8 << 16 | feature(right) << 8 | feature(left)
    //     An internal node names the features its two children test instead of its own (a leaf child
    //     counts as feature 0), so a walk fetches the next node and the next feature value at the same
    //     time: one shared-memory round trip per level instead of two. Needs n_features <= 256 (cf_ok);
    //     missing_go_left is read from d_nodes on the (rare) NaN path.
    uint2* d_nodes_f0 = nullptr;     // [n_nodes] own-feature
    uint2* d_nodes_f1 = nullptr;     // [n_nodes] child-feature
    uint8_t* d_rootfeat = nullptr;   // [n_trees] feature tested by the root (0 for a single-leaf tree)
    bool fused_ok = false, cf_ok = false;
    int32_t max_depth = 0;
    std::vector<int64_t> h_node_offset;   // host copy, for building group tables
    // tree groups staged into shared memory by the fused kernel, one table per
    // (buffer nodes, chunk) configuration, built on first use (pk_forest_groups)
    struct GroupTable { int tbn = 0, chunk = 0; int4* d = nullptr; int32_t n = 0; };
    std::vector<GroupTable> group_tables;
};

// Group table for buffers of `tbn` nodes: consecutive trees whose nodes fit one buffer
// (trimmed to a multiple of `chunk` trees when more than one chunk fits); a tree larger
// than the buffer is its own group and only its first `tbn` nodes are staged.
//   .x first tree, .y number of trees, .z first staged node (even),
//   .w staged nodes (even); negative when the group's single tree is only partly staged
int pk_forest_groups(pk_forest* f, int tbn, int chunk, const int4** d_groups, int32_t* n_groups);

struct pk_chrom {
    int device = 0;
    cudaStream_t stream = nullptr;
    int32_t n = 0, w = 0, S = 0, F = 0;
    int32_t lower = 0, upper = 0;    // effective (clamped)
    int32_t ND = 0;                  // stored diagonals = upper + 2w + 1 = exp_len
    int64_t pitch = 0;               // band row pitch (elements)
    int balanced = 0;
    // state
    bool has_pixels = false, has_expected = false, has_candidates = false, has_scores = false;
    int32_t row_begin = 0, row_end = 0;
    bool whole = true;
    // device buffers
    int32_t* d_band = nullptr;       // [ND][pitch] raw counts, diagonal-major
    // Row-major copy of the band for the fused kernel's window fetch: band2[r * P2 + o] = count of pixel (r, r + o),
    // o < ND - 1 (the trimmed diagonal and the padding columns are zero). P2 = 4k + 1, so that the skewed tensor map
    // T[j][i] = band2 + j * (P2 - 1) + i (row stride a multiple of 16 bytes) addresses the dense matrix cell
    // (row j, column i) and a pixel's (2w+1)^2 window is one rectangular TMA box (pk_fused.cu, phase A).
    int32_t* d_band2 = nullptr;      // [n][P2] (+ padding)
    int64_t P2 = 0;
    bool band2_valid = false;        // built from the current pixels
    alignas(64) CUtensorMap tmap;    // the skewed view, box = (2w+1) rows x (2w+1 + slack) columns
    bool tmap_ok = false;
    double* d_w = nullptr;           // [n]
    double* d_wp = nullptr;          // [n] weights of the Poisson filter when they differ from d_w (pk_chrom_set_poisson_weights)
    bool use_wp = false;
    uint8_t* d_valid = nullptr;      // [n]
    uint32_t* d_vbits = nullptr;     // [ceil(n/32)] the same, one bit per bin
    double* d_scratch = nullptr;     // [ND][pitch] compacted diagonal values
    double* d_diag_sum = nullptr;    // [ND]
    long long* d_diag_cnt = nullptr; // [ND]
    double* d_exp = nullptr;         // [ND]
    double* d_bg = nullptr;          // [ND]
    // [4]: 0 = a count exceeded the Poisson table, 1 = largest in-band count,
    //      2 = bit0 expected fit failed / bit1 diagonal too long / bit2 a weight outside [1e-45, 1e45],
    //      3 = bit0 pixels not sorted / bit1 candidate buffer too small
    int32_t* d_flags = nullptr;
    unsigned char* d_head = nullptr; // flags | d_ncand | d_counters | d_batch_win in one block (one copy reads them all)
    size_t head_bytes = 0;
    // upload staging
    int32_t *d_b1 = nullptr, *d_b2 = nullptr, *d_cnt = nullptr;
    int64_t pix_cap = 0;
    long long* d_rowptr = nullptr;   // [n+1]
    bool declared_sorted = false;
    unsigned char* d_blob = nullptr; // packed pixel rows (pk_chrom_upload_rows), or the caller's device blob
    unsigned char* d_blob_own = nullptr;   // the staging block this handle owns
    int64_t blob_cap = 0;
    long long rows_hdr[16] = {};     // host copy of the blob's header
    // candidates
    int32_t n_chunks = 0;
    unsigned long long* d_cstate = nullptr;   // [nd_cand * n_chunks] candidate counts per scan tile (all rows, row tile) as uint2
    uint32_t* d_bits = nullptr;               // [nd_cand * n_chunks * PK_CTILE / 32] one bit per band slot
    int64_t cnt_cap = 0;
    long long* d_ncand = nullptr;    // [2]: candidates in the row tile, in the whole chromosome
    int64_t n_cand = 0, n_cand_all = 0;   // host copies, valid when n_cand_known
    bool n_cand_known = false;
    int64_t cand_cap = 0;
    // last scoring request (replayed if a device-side capacity flag was raised)
    pk_forest* last_forest = nullptr;
    double last_thre = 0.0;
    int32_t *d_cx = nullptr, *d_cd = nullptr, *d_crank = nullptr;
    // scoring
    uint8_t* d_keep = nullptr;
    float* d_fea32 = nullptr;        // [n_cand][F]
    int64_t fea_cap = 0, keep_cap = 0;
    double* d_prob = nullptr;
    int32_t* d_batch_win = nullptr;  // [n_batches]
    int64_t n_batches = 0, batch_cap = 0;
    // records
    int32_t *d_rx = nullptr, *d_ry = nullptr, *d_rb = nullptr;
    double *d_rp = nullptr, *d_rv = nullptr;
    unsigned long long* d_counters = nullptr;   // [4]: 0 = n_records, 1 = n_windows
    int64_t rec_cap = 0;
    // records sorted by (x, y) and packed right after the scoring pass (no kernel at fetch time)
    int32_t *d_rowcnt = nullptr, *d_rowoff = nullptr, *d_rrank = nullptr;   // [n], [n+1], [cand_cap]
    uint32_t* d_perm = nullptr;         // [eager_cap]
    unsigned char* d_packed = nullptr;
    int64_t eager_cap = 0;              // records the eager ordering covers
    int64_t rrank_cap = 0;
    bool eager_valid = false;           // d_packed belongs to the current scores
    // pixel columns of the last upload as they sit on the device (pk_chrom_depth): kind 0 none,
    // 1 COO (b1, b2, cnt int32), 2 rows + int32 columns, 3 rows + uint16 (bin2 - bin1, count), 4 packed rows (d_blob)
    int up_kind = 0;
    const void *up_b1 = nullptr, *up_b2 = nullptr, *up_cnt = nullptr;
    const long long* up_rowptr = nullptr;
    int64_t up_nnz = 0;
    cudaStream_t score_stream = nullptr;   // optional second stream for the scoring pass
    bool use_score_stream = false;
    cudaEvent_t ev_x = nullptr;            // hand-over between the two streams
    bool reuse_pending = false;            // engine: ev_x marks the end of the previous unit's record copy
    unsigned long long h_counts[4] = {0, 0, 0, 0};
    bool counts_valid = false;          // h_counts read since the last scoring pass
    unsigned char* h_stage = nullptr;   // pinned staging for fetch_results
    size_t h_stage_bytes = 0;
    // timing
    cudaEvent_t ev[16] = {};
    float stage_ms[8] = {};
    bool timing = true;                 // record the per-stage events (pk_chrom_stage_ms); the engine turns it off
};

// Poisson decision table (host long double -> float64), per device copy
int pk_poisson_table_host(int32_t k_max, const double** out);      // grows a process-wide table
int pk_poisson_table_device(int device, int32_t k_min_size, const double** d_out, int32_t* k_max_out);

// expected-curve fit on the host (PAVA + numpy.interp restatement)
int pk_fit_expected_host(const double* sum, const long long* cnt, int32_t len, double* out_exp);
