/*
 * peakachu_b200 -- C ABI of the B200 (sm_100a) loop-scoring path.
 *
 * The reference (tariks/peakachu v2.3) is pure Python and has no FFI. The seam this
 * library sits behind is the class peakachu/scoreUtils.py:9-135 (`Chromosome`) as
 * driven by peakachu/score_chromosome.py:3-71 and peakachu/score_genome.py:3-84.
 * Each entry point below names the reference code it replaces. The binding a
 * maintainer adds on the reference side is a ctypes stub; see INTEGRATION.md.
 *
 * Conventions
 *  - every call returns 0 on success, a negative PK_E* code on failure;
 *    pk_last_error() returns a thread-local message for the last failure;
 *    no C++ exception crosses this boundary.
 *  - the caller owns every buffer it passes in; the library owns the opaque
 *    handles it returns until *_destroy.
 *  - a handle is bound to one CUDA device; all work of a handle is issued on the
 *    stream given at creation (a cudaStream_t passed as void*, NULL = default
 *    stream). One host thread per handle at a time.
 *  - `mem` arguments say where a caller buffer lives: PK_MEM_HOST or PK_MEM_DEVICE.
 *  - calls are asynchronous on the handle's stream unless they return data to
 *    host memory (those synchronise the stream).
 */
#ifndef PEAKACHU_B200_H
#define PEAKACHU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_OK            0
#define PK_EINVAL       -1   /* bad argument */
#define PK_ECUDA        -2   /* CUDA runtime error (message has the cudaError string) */
#define PK_ENOMEM       -3
#define PK_ESTATE       -4   /* call out of order (e.g. score before set_expected) */
#define PK_ECAPACITY    -5   /* caller buffer too small */
#define PK_EUNSUPPORTED -6

#define PK_MEM_HOST   0
#define PK_MEM_DEVICE 1
/* OR-ed into `mem` of pk_chrom_upload_pixels: the pixels are in cooler order (sorted by
 * bin1, then bin2, bin1 <= bin2). Enables the tiled band build; verified on the device. */
#define PK_PIXELS_SORTED 0x100

typedef struct pk_forest pk_forest;
typedef struct pk_chrom pk_chrom;

const char* pk_last_error(void);
int pk_abi_version(void);
/* number of visible CUDA devices; fails with PK_ECUDA when there is none */
int pk_device_count(int* out);
/* PCI bus id of a device ("0000:1b:00.0"), NUL-terminated into buf[len]. The host side uses it to run each
 * rank of score_genome (score_genome.py:46-84, one process per GPU here) on the NUMA node its GPU hangs off,
 * so that the pinned pixel columns are local to the DMA engine. */
int pk_device_pci_bus_id(int device, char* buf, int len);

/* ---- forest: replaces model.predict_proba(fea)[:, 1] (scoreUtils.py:109) on the
 * joblib-loaded sklearn RandomForestClassifier (score_chromosome.py:14). Arrays are
 * the concatenated per-tree sklearn `tree_` tables (host memory): children are
 * tree-local indices, -1 for leaves; `leaf_p1[i]` is the class-1 fraction
 * tree_.value[i,0,1]. Traversal and accumulation order follow sklearn
 * (_tree.pyx::_apply_dense, _forest.py predict_proba): float32 feature <= float64
 * threshold goes left, NaN follows missing_go_to_left, float64 sum in tree order,
 * one divide by n_trees. */
int pk_forest_create(int device, int32_t n_trees, int32_t n_features,
                     const int64_t* node_offset /* [n_trees+1] */,
                     const int32_t* feature, const double* threshold,
                     const int32_t* left, const int32_t* right,
                     const uint8_t* missing_left, const double* leaf_p1,
                     pk_forest** out);
int pk_forest_destroy(pk_forest* f);
int pk_forest_info(const pk_forest* f, int32_t* n_trees, int32_t* n_features, int64_t* n_nodes);
/* parity tap: leaves (tree-local sklearn node id, as estimator.apply) and/or
 * probabilities for caller-supplied float32 rows. X, leaves, proba are DEVICE
 * pointers; leaves is [n_rows][n_trees] or NULL; proba is [n_rows] or NULL. */
int pk_forest_apply(pk_forest* f, const float* X, int64_t n_rows,
                    int32_t* leaves, double* proba, void* stream);

/* ---- chromosome: replaces scoreUtils.Chromosome (scoreUtils.py:9-135).
 * lower/upper are the user's -l/-u; the library clamps them like scoreUtils.py:13-14
 * (lower = max(lower, w+1), upper = min(upper, n-2w)). balanced != 0 is the
 * `--clr-weight-name <col>` mode (weights given), 0 is `--clr-weight-name raw`. */
int pk_chrom_create(int device, int32_t n_bins, int32_t width, int32_t lower, int32_t upper,
                    int balanced, void* stream, pk_chrom** out);
int pk_chrom_destroy(pk_chrom* c);
/* effective (clamped) bounds and the expected-curve length upper+2w+1 */
int pk_chrom_bounds(const pk_chrom* c, int32_t* lower_eff, int32_t* upper_eff, int32_t* exp_len);

/* Upper-triangle pixels of one chromosome (bin1 <= bin2, chromosome-local ids, any
 * order, duplicates summed as utils.tocsr does, utils.py:10-15) plus the balancing
 * weight per bin (NULL in raw mode). Builds the dense diagonal-major count band,
 * the per-bin `valid` mask of utils.calculate_expected (utils.py:146-156) and the
 * per-diagonal sums (utils.py:160-170, numpy pairwise order). Replaces the
 * cooler fetches + tocsr + band trim (score_chromosome.py:42-44, scoreUtils.py:30-33). */
int pk_chrom_upload_pixels(pk_chrom* c, const int32_t* bin1, const int32_t* bin2,
                           const int32_t* count, int64_t nnz, const double* weights, int mem);
/* Same from cooler's own CSR layout: bin1_offset[n_bins+1] (indexes/bin1_offset restricted
 * to the chromosome, rebased to 0) plus the bin2 / count columns of those pixels. */
int pk_chrom_upload_csr(pk_chrom* c, const int64_t* bin1_offset, const int32_t* bin2, const int32_t* count,
                        int64_t nnz, const double* weights, int mem);
/* Same with narrow columns: bin2_delta[p] = bin2 - bin1 and count[p], both uint16 -- 4 bytes per
 * pixel cross the bus instead of 8. Every pixel of the chromosome must be representable
 * (bin2 - bin1 <= 65535, count <= 65535); a caller that drops farther pixels instead must make
 * sure no bin loses its last finite pixel, because the `valid` mask of utils.py:146-156 looks
 * at the whole matrix. */
int pk_chrom_upload_csr16(pk_chrom* c, const int64_t* bin1_offset, const uint16_t* bin2_delta, const uint16_t* count,
                          int64_t nnz, const double* weights, int mem);
/* Packed pixel rows: the whole chromosome in ONE contiguous blob, about 1.3 bytes per band pixel
 * on a dense map (peakachu_b200/rowpack.py writes and reads it; coolio.PKCool stores it). Little
 * endian, every section 16-byte aligned, offsets from the start of the blob:
 *   header  int64[16]: PK_ROWS_MAGIC, n_bins, nd_enc, words_per_row = ceil(nd_enc / 32), nnz_band, n_esc,
 *                      n_far, off_bits, off_cnt_off, off_cnt8, off_esc, off_far_off, off_far_b2,
 *                      off_far_cnt, total_bytes, 0
 *   bits    uint32[n_bins][words_per_row]  bit d of row x: pixel (x, x + d) is present (d < nd_enc)
 *   cnt_off uint32[n_bins + 1]             first count byte of each row
 *   cnt8    uint8[nnz_band]                counts in (row, distance) order; 255 = see `esc`
 *   esc     int32[3][n_esc]                x | d | count of the pixels with count >= 255
 *   far_off int64[n_bins + 1], far_b2 int32[n_far], far_cnt int32[n_far]
 *                                          pixels with d >= nd_enc as CSR columns: they never enter the
 *                                          band but decide the `valid` mask (utils.py:146-156) and `depth`
 * nd_enc must cover the band: nd_enc >= upper + 2w + 1 (effective upper). No duplicates by construction. */
#define PK_ROWS_MAGIC 0x31524B50LL
int pk_chrom_upload_rows(pk_chrom* c, const void* blob, int64_t bytes, const double* weights, int mem);
/* Balancing weights of the Poisson filter (scoreUtils.py:55-57) when they differ from the ones the pixel
 * values were balanced with. cooler inverts a weight column whose `divisive_weights` attribute is set
 * (hic2cool's KR / VC columns) inside matrix(balance=name), whereas the reference passes the column's raw
 * values to Chromosome (score_chromosome.py:44): such a map is uploaded with 1 / w and gets the raw column
 * here. Call after the upload, before pk_chrom_find_candidates. */
int pk_chrom_set_poisson_weights(pk_chrom* c, const double* weights, int mem);
/* `peakachu depth` (calculate_depth.py:25-28): sum of the raw counts of the pixels of the last
 * upload with bin2 - bin1 >= min_dis_bins. Columns passed as device pointers must still be alive. */
int pk_chrom_depth(pk_chrom* c, int32_t min_dis_bins, int64_t* total);
/* per-diagonal (sum, n_valid) for d = 0..upper+2w, to HOST arrays of exp_len */
int pk_chrom_diag_sums(pk_chrom* c, double* out_sum, int64_t* out_cnt);
/* expected curve fitted by the library itself: mean where n_valid > 10, then the
 * non-increasing isotonic fit + clipped linear interpolation of utils.py:173-176
 * (the library's own PAVA + interpolation; scikit-learn is not involved). */
int pk_chrom_fit_expected(pk_chrom* c);
/* or supply it (HOST arrays of exp_len): exp_arr normalises windows, background
 * drives the Poisson filter (scoreUtils.py:16-24) */
int pk_chrom_set_expected(pk_chrom* c, const double* exp_arr, const double* background);
int pk_chrom_get_expected(pk_chrom* c, double* out_exp /* HOST [exp_len] */);

/* Candidate selection, scoreUtils.py:40-68: raw count > 0 and Poisson upper tail
 * < 0.01 against background[d] / (w_x * w_y). Restricted to rows x in
 * [row_begin, row_end) (pass 0, n_bins for the whole chromosome) -- the band
 * row-tile seam used for multi-GPU sharding. n_candidates (HOST) may be NULL. */
int pk_chrom_find_candidates(pk_chrom* c, int32_t row_begin, int32_t row_end, int64_t* n_candidates);
/* parity taps (HOST outputs, reference order: distance asc, row asc) */
int pk_chrom_candidates(pk_chrom* c, int32_t* out_x, int32_t* out_y, int64_t capacity, int64_t* n);
/* getwindow tap (scoreUtils.py:70-93) over the current candidates: keep[i] = window
 * passed the border mask and utils.distance_normalize's filters; fea32/fea64 are
 * [n_candidates][(2w+1)^2] HOST arrays (either may be NULL), rows of rejected
 * candidates are left untouched. */
int pk_chrom_features(pk_chrom* c, uint8_t* keep, float* fea32, double* fea64, int64_t capacity);
/* The same tap taken INSIDE the product kernel: pk_chrom_score's fused kernel builds the float32
 * feature rows in shared memory and normally never writes them out; here it also spills them to
 * global memory (what sklearn's predict_proba would be handed, scoreUtils.py:109). keep / fea32 are
 * HOST arrays as above; rows of rejected candidates come back zero. PK_EUNSUPPORTED for shapes the
 * fused kernel does not cover. Scores of an earlier pk_chrom_score are invalidated. */
int pk_chrom_fused_features(pk_chrom* c, pk_forest* f, uint8_t* keep, float* fea32, int64_t capacity);
/* The same windows at caller-supplied pixels (x[i], y[i]), x <= y, HOST int32 arrays: the
 * training-set extraction of trainUtils.buildmatrix (trainUtils.py:12-44; the caller applies
 * its coordinate mask, trainUtils.py:22). Needs pixels and an expected curve covering
 * max(y - x) + 2w + 1 distances. The handle's candidate list is replaced: call
 * pk_chrom_find_candidates again before scoring. keep / fea32 / fea64 as above, [n] rows. */
int pk_chrom_features_at(pk_chrom* c, const int32_t* x, const int32_t* y, int64_t n, uint8_t* keep, float* fea32,
                         double* fea64);

/* Chromosome.score (scoreUtils.py:95-125): features -> forest -> keep prob > min_prob
 * (strict). The reference walks candidates in batches of 100,000 (in its own order,
 * over the whole chromosome) and silently drops a batch in which <= 1 window
 * survives the filters (scoreUtils.py:104-108). Every candidate carries its
 * whole-chromosome rank, so batch ids are right for row tiles too; when the handle
 * covers the whole chromosome the drop rule is applied on the device, for a row
 * tile the caller applies it after summing pk_chrom_batch_windows over the tiles.
 * Records stay on the device until fetched. */
int pk_chrom_score(pk_chrom* c, pk_forest* f, double min_prob);
int pk_chrom_result_count(pk_chrom* c, int64_t* n_records, int64_t* n_candidates, int64_t* n_windows);
/* surviving windows per reference batch (HOST array; n_batches may exceed capacity -> PK_ECAPACITY) */
int pk_chrom_batch_windows(pk_chrom* c, int64_t* out, int64_t capacity, int64_t* n_batches);
/* records (x, y, prob, balanced value[, batch id]); out_batch may be NULL.
 * mem = PK_MEM_HOST: sorted by (x, y) like prob_csr.nonzero() (scoreUtils.py:130).
 * mem = PK_MEM_DEVICE: device-to-device copy in emission order (unsorted). */
int pk_chrom_fetch_results(pk_chrom* c, int32_t* out_x, int32_t* out_y, double* out_prob,
                           double* out_val, int32_t* out_batch, int64_t capacity, int mem);

/* ---- engine: the chromosome loop of score_genome.main (score_genome.py:46-84) as a persistent
 * per-device pipeline. Streams, chromosome handles and pinned staging are created once and reused;
 * pk_engine_submit queues one unit -- a chromosome, or rows [row_begin, row_end) of one (the
 * multi-GPU row-tile seam) -- and never waits for the device; pk_engine_collect waits for all queued
 * units, in submission order, and exposes their records. One host thread per engine.
 * Unit columns are HOST arrays (pinned memory makes the uploads asynchronous); they must stay alive
 * until pk_engine_collect returns. */
typedef struct pk_engine pk_engine;
#define PK_ENC_COO   0   /* a = bin1 int32[size], b = bin2 int32[size], c = count int32[size], cooler order */
#define PK_ENC_CSR32 1   /* a = bin1_offset int64[n_bins+1], b = bin2 int32[size], c = count int32[size] */
#define PK_ENC_CSR16 2   /* a = bin1_offset int64[n_bins+1], b = (bin2 - bin1) uint16[size], c = count uint16[size] */
#define PK_ENC_ROWS  3   /* a = packed rows blob (pk_chrom_upload_rows), size = its bytes */
typedef struct pk_unit {
    int64_t tag;                 /* returned with the results */
    int32_t n_bins, row_begin, row_end, encoding;
    const void *a, *b, *c;
    int64_t size;
    const double* weights;       /* NULL: raw mode */
    const double* poisson_weights;   /* NULL: the same as `weights` (see pk_chrom_set_poisson_weights) */
    double min_prob;
} pk_unit;
typedef struct pk_unit_result {
    int64_t tag;
    int32_t n_bins, row_begin, row_end, whole;
    int64_t n_records, n_candidates, n_windows, n_batches;
    /* records ordered by (x, y); pinned memory owned by the engine, valid until the next pk_engine_submit */
    const int32_t *x, *y, *batch;
    const double *prob, *value;
    const int32_t* batch_windows;   /* [n_batches] surviving windows per reference batch (see pk_chrom_score) */
} pk_unit_result;
/* depth: chromosomes in flight (uploads of the next ones overlap the kernels of the current one) */
int pk_engine_create(int device, pk_forest* f, int32_t width, int32_t lower, int32_t upper, int depth, pk_engine** out);
int pk_engine_destroy(pk_engine* e);
int pk_engine_submit(pk_engine* e, const pk_unit* u);
/* out == NULL: only the number of queued units. */
int pk_engine_collect(pk_engine* e, pk_unit_result* out, int64_t capacity, int64_t* n_units);
/* after a failed submit / collect: wait for the device and drop the queued units */
int pk_engine_reset(pk_engine* e);

/* ---- Poisson decision table: crit[k] = smallest float64 mu with
 * Pr[Poisson(mu) > k] >= 0.01, so that `p < 0.01` <=> `mu < crit[k]`
 * (scipy.stats.poisson.sf is increasing in mu). HOST output, for tests. */
int pk_poisson_critical_mu(int32_t k_max, double* out /* [k_max+1] */);

/* the expected-curve fit alone on HOST arrays (test hook for the restated
 * IsotonicRegression(increasing=False, out_of_bounds='clip') of utils.py:173-176) */
int pk_fit_expected(const double* sum, const int64_t* cnt, int32_t len, double* out_exp);

/* device self-test: the reciprocal-based division used by the feature kernel against IEEE
 * division on n pseudo-random / adversarial operand pairs; *mismatches must come back 0 */
int pk_selftest_divide(int device, int64_t n, uint64_t seed, int64_t* mismatches);

/* Chromosome.writeBed (scoreUtils.py:127-135) for n records: tab-separated
 * chrom, x*res, (x+1)*res, chrom, y*res, (y+1)*res, str(prob), str(value) with floats in
 * Python's shortest round-trip repr. HOST buffers; *written = bytes produced (or needed). */
int pk_format_bedpe(const char* chrom, int64_t res, const int32_t* x, const int32_t* y, const double* prob,
                    const double* val, int64_t n, char* out, int64_t capacity, int64_t* written);

/* Host-only helper of the built-in .cool reader (peakachu_b200/h5mini.py; cooler's own reader is h5py / libhdf5,
 * score_chromosome.py:33-43): decode the chunks of a 1-D HDF5 dataset that cover elements [lo, hi) straight from the
 * memory-mapped file into `out` ([hi - lo] elements) on n_threads host threads (0: all). chunk_off / chunk_bytes locate
 * the filtered chunks in `file`, first_elem is each chunk's first element; filters as flags: deflate (zlib), shuffle,
 * fletcher32 (its 4 trailing bytes are dropped, not verified). */
int pk_h5_decode_chunks(const uint8_t* file, int64_t n_chunks, const int64_t* chunk_off, const int64_t* chunk_bytes,
                        const int64_t* first_elem, int64_t chunk_elems, int32_t elem_size, int32_t deflate,
                        int32_t shuffle, int32_t fletcher32, int64_t lo, int64_t hi, void* out, int32_t n_threads);

/* Host-only: packed pixel rows (the blob of pk_chrom_upload_rows, byte for byte what peakachu_b200/rowpack.py writes)
 * from the rows of ONE chromosome as a cooler file stores them (score_chromosome.py:42-43 fetches this block through
 * cooler): bin1_offset int64[n_bins + 1] relative to the first pixel handed in, bin2 as int32 / int64 (bin2_bytes 4 | 8)
 * genome-wide ids of which bin2_base (the chromosome's first bin) is subtracted -- pixels behind the chromosome are
 * inter-chromosomal and dropped --, count as int32 / int64 / float64 (count_kind 0 | 1 | 2). Duplicates are summed and
 * zero counts dropped as utils.tocsr would (utils.py:10-15). far_mode 0 keeps every pixel with bin2 - bin1 >= nd_enc in
 * the far lists (what `depth` needs); far_mode 1 keeps only those the scoring path needs: a far pixel never enters the
 * band, it only makes its two bins `valid` (utils.py:146-156), so it is dropped unless it is finite -- count > 0 and,
 * with `weights` (float64[n_bins], the balancing weights; NULL: raw counts), isfinite((w_x w_y) count) -- and one of
 * its bins has no finite pixel inside nd_enc; the `valid` mask the device derives is the same, and a deep genome-wide
 * map sheds the bulk of its bytes. Call with out = NULL to learn the size (*needed), then with a buffer of at least
 * that many bytes. PK_EINVAL names the first offending row: pixels below the diagonal, out of order, negative,
 * beyond int32 or fractional counts. n_threads 0: all host threads. */
int pk_rows_pack(const int64_t* bin1_offset, const void* bin2, int32_t bin2_bytes, int64_t bin2_base, const void* count,
                 int32_t count_kind, int64_t n_bins, int32_t nd_enc, int32_t far_mode, const double* weights, void* out,
                 int64_t capacity, int64_t* needed, int32_t n_threads);

/* a non-blocking CUDA stream for pk_chrom_create, for callers that do not bring their own
 * (two handles on two streams overlap one chromosome's upload with another's kernels) */
int pk_stream_create(int device, void** out);
int pk_stream_destroy(int device, void* stream);
/* priority > 0: kernels of this stream are scheduled ahead of those of ordinary streams */
int pk_stream_create_priority(int device, int priority, void** out);
/* Run the fused scoring kernel of this handle (pk_chrom_score) on `stream` instead of the
 * handle's own; the handle's stream waits for it. With the handles' own streams at high
 * priority and one ordinary stream shared by the scoring kernels, those run back to back and the
 * short stages of the chromosomes queued behind slip in where one ends and the next starts.
 * NULL: back to one stream. */
int pk_chrom_set_score_stream(pk_chrom* c, void* stream);

/* return the library's cached device blocks (of destroyed handles) to the driver */
int pk_release_memory(void);

/* process-wide tuning knobs, for benchmarking and A/B runs (results are identical under every setting):
 *   "fused"           -1 auto (default), 0 separate feature + forest kernels, 1 + v: fused-kernel variant v
 *                     (tile sizes, warp groups; the list is in pk_fused.cu, pk_launch_fused)
 *   "prune"           1 (default): the fused kernel stops walking trees for pixels whose probability can no
 *                     longer exceed min_prob (emitted records are unaffected); 0 walks every tree
 *   "child_features"  forest walk on the node encoding that names the children's features: -1 where measured
 *                     faster (w = 7), 0 off, 1 on
 *   "tma"             windows fetched as TMA boxes (cp.async.bulk.tensor.2d) from a row-major copy of the band:
 *                     1 (default) where measured faster (w = 7), 2 always (w = 5 too), 0 never. Read when a
 *                     chromosome handle is created (the copy is allocated there).
 *   "reserve_sms"     SMs the fused kernel leaves free for the short stages of other chromosomes in flight */
int pk_set_tuning(const char* key, int value);

/* time spent (ms, CUDA events) in each stage of the last upload/fit/find/score of
 * this handle: [0] band build, [1] diagonal sums, [2] expected fit, [3] candidate
 * scan, [4] window features (or the fused features+forest kernel, then [5] = 0),
 * [5] forest, [6] emit. HOST array of 8. */
int pk_chrom_stage_ms(pk_chrom* c, float* out_ms);

#ifdef __cplusplus
}
#endif
#endif /* PEAKACHU_B200_H */
