// peakachu_b200: C ABI (include/peakachu_b200.h) -- handle management, forest
// packing, stage orchestration. No kernels here; see pk_kernels.cu.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "pk_common.cuh"

// launchers implemented in pk_kernels.cu / pk_stages.cu / pk_fused.cu
int pk_launch_scatter(pk_chrom* c, const int32_t* b1, const int32_t* b2, const int32_t* cnt, int64_t nnz);
int pk_launch_band_csr(pk_chrom* c, const long long* rowptr, const void* b2, const void* cnt, int enc);
int pk_launch_band_rows(pk_chrom* c);
int pk_launch_band_rowmajor(pk_chrom* c);
int pk_launch_rowptr(pk_chrom* c, const int32_t* b1, const int32_t* b2, int64_t nnz, long long* rowptr);
int pk_launch_diag_sums(pk_chrom* c);
bool pk_fit_on_device_supported(int len);
int pk_launch_fit_expected(pk_chrom* c);
int pk_launch_candidates(pk_chrom* c, const double* d_crit, int kmax);
int pk_launch_features(pk_chrom* c, double* d_fea64);
int pk_launch_forest(const pk_forest* f, const float* X, const uint8_t* keep, int64_t n_rows, int32_t* leaves,
                     double* proba, cudaStream_t stream);
int pk_launch_emit(pk_chrom* c, double thre);
int pk_launch_fused(pk_chrom* c, pk_forest* f, int variant, double thre, int reserve_sms, int child_features, int tma, float* fea_tap);
int pk_launch_sort_records_eager(pk_chrom* c, long long M);
int pk_launch_depth(pk_chrom* c, int32_t min_dis, unsigned long long* d_total);
bool pk_fused_supported(int w, int n_trees);

int pk_run_selftest_divide(long long n, unsigned long long seed, long long* mismatches);

extern "C" int pk_selftest_divide(int device, int64_t n, uint64_t seed, int64_t* mismatches) {
    if (!mismatches || n < 0) { pk_set_error("pk_selftest_divide: bad argument"); return PK_EINVAL; }
    PK_CUDA(cudaSetDevice(device));
    long long m = 0;
    PK_CHECK(pk_run_selftest_divide(n, seed, &m));
    *mismatches = m;
    return PK_OK;
}

// tuning knobs (pk_set_tuning): fused = -1 auto, 0 unfused kernels, 1 + v: fused kernel variant v
static int g_tune_fused = -1;
static int g_tune_prune = 1;
static int g_tune_cf = -1;       // fused forest walk on the child-feature node encoding: -1 where measured faster (w = 7), 0 off, 1 on
// fused kernel fetches windows as TMA boxes from a row-major copy of the band: 1 = where measured faster (w = 7: -8 %;
// the w = 5 kernel is 6 % slower with it, profiles/r2_summary.md), 2 = always, 0 = never (per-cell gather)
static int g_tune_tma = 1;
static int g_tune_reserve = 8;   // SMs the fused kernel leaves to the short stages of other chromosomes (pipelined use: handles with a
                                 // score stream; measured on the c2 chromosome end to end: 0.663 ms with 0, 0.633-0.641 with 4-12, profiles/r2_summary.md)     // retire pixels that cannot exceed min_prob (exact for every emitted record)

extern "C" int pk_set_tuning(const char* key, int value) {
    if (key && !strcmp(key, "fused")) { g_tune_fused = value; return PK_OK; }
    if (key && !strcmp(key, "prune")) { g_tune_prune = value; return PK_OK; }
    if (key && !strcmp(key, "child_features")) { g_tune_cf = value; return PK_OK; }
    if (key && !strcmp(key, "tma")) { g_tune_tma = value; return PK_OK; }
    if (key && !strcmp(key, "reserve_sms")) { g_tune_reserve = value < 0 ? 0 : value; return PK_OK; }
    pk_set_error("pk_set_tuning: unknown key %s", key ? key : "(null)");
    return PK_EINVAL;
}

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

void pk_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

extern "C" const char* pk_last_error(void) { return g_err.c_str(); }
extern "C" int pk_abi_version(void) { return 1; }

extern "C" int pk_device_count(int* out) {
    if (!out) { pk_set_error("pk_device_count: out is NULL"); return PK_EINVAL; }
    int n = 0;
    PK_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0) { pk_set_error("no CUDA device visible"); return PK_ECUDA; }
    *out = n;
    return PK_OK;
}

extern "C" int pk_device_pci_bus_id(int device, char* buf, int len) {
    if (!buf || len < 13) { pk_set_error("pk_device_pci_bus_id: buffer of at least 13 bytes needed"); return PK_EINVAL; }
    PK_CUDA(cudaDeviceGetPCIBusId(buf, len, device));
    return PK_OK;
}

// ---------------------------------------------------------------------------
// device memory: a small caching allocator. Handles are created and destroyed per
// chromosome (as the reference builds one Chromosome object per chromosome), so
// cudaMalloc/cudaFree -- both synchronising -- must not sit on that path. Freed
// blocks are kept per device and reused for requests of up to twice their size.
// ---------------------------------------------------------------------------
namespace {
std::mutex g_pool_mu;
std::map<int, std::multimap<size_t, void*>> g_pool_free;   // device -> cached blocks by size
std::map<void*, std::pair<int, size_t>> g_pool_live;       // pointer -> (device, bytes)
}  // namespace

static int pool_alloc(void** out, size_t bytes) {
    *out = nullptr;
    if (bytes == 0) bytes = 256;
    bytes = (bytes + 255) & ~(size_t)255;
    int dev = 0;
    PK_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto& fl = g_pool_free[dev];
        auto it = fl.lower_bound(bytes);                 // smallest cached block that is large enough
        if (it != fl.end() && it->first <= 2 * bytes + 4096) {
            *out = it->second;
            g_pool_live[*out] = {dev, it->first};
            fl.erase(it);
            return PK_OK;
        }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {
        pk_set_error("cudaMalloc(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? PK_ENOMEM : PK_ECUDA;
    }
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool_live[*out] = {dev, bytes};
    return PK_OK;
}

// INVARIANT: a block is freed only after the stream that used it has been synchronised
// (pk_chrom_destroy, settle_candidates / read_flags before a regrow, pk_chrom_fetch_results): a cached
// block carries no stream ordering and may be handed to a handle on another stream at once. Buffers
// that grow while work may still be queued call quiesce() first.
static void pool_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool_live.find(p);
    if (it == g_pool_live.end()) { cudaFree(p); return; }
    g_pool_free[it->second.first].emplace(it->second.second, p);
    g_pool_live.erase(it);
}

// pinned host staging buffers are process-wide and reused (cudaMallocHost / cudaFreeHost
// synchronise the device)
namespace {
std::mutex g_hstage_mu;
std::vector<std::pair<unsigned char*, size_t>> g_hstage_free;
}  // namespace

static int hstage_acquire(unsigned char** out, size_t* out_bytes, size_t bytes) {
    {
        std::lock_guard<std::mutex> lk(g_hstage_mu);
        for (size_t i = 0; i < g_hstage_free.size(); ++i)
            if (g_hstage_free[i].second >= bytes) {
                *out = g_hstage_free[i].first; *out_bytes = g_hstage_free[i].second;
                g_hstage_free.erase(g_hstage_free.begin() + i);
                return PK_OK;
            }
    }
    const size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMallocHost((void**)out, want);
    if (e != cudaSuccess) { pk_set_error("cudaMallocHost(%zu): %s", want, cudaGetErrorString(e)); return PK_ENOMEM; }
    *out_bytes = want;
    return PK_OK;
}

static void hstage_release(unsigned char* p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_hstage_mu);
    g_hstage_free.push_back({p, bytes});
}

extern "C" int pk_stream_create(int device, void** out) {
    if (!out) { pk_set_error("pk_stream_create: out is NULL"); return PK_EINVAL; }
    PK_CUDA(cudaSetDevice(device));
    cudaStream_t s = nullptr;
    PK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = (void*)s;
    return PK_OK;
}

// priority > 0: the stream's kernels are scheduled ahead of those of ordinary streams
extern "C" int pk_stream_create_priority(int device, int priority, void** out) {
    if (!out) { pk_set_error("pk_stream_create_priority: out is NULL"); return PK_EINVAL; }
    PK_CUDA(cudaSetDevice(device));
    int lo = 0, hi = 0;                       // numerically lower = higher priority
    PK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s = nullptr;
    PK_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, priority > 0 ? hi : lo));
    *out = (void*)s;
    return PK_OK;
}

extern "C" int pk_stream_destroy(int device, void* stream) {
    if (!stream) return PK_OK;
    PK_CUDA(cudaSetDevice(device));
    PK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    PK_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return PK_OK;
}

extern "C" int pk_release_memory(void) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& kv : g_pool_free) {
        cudaSetDevice(kv.first);
        cudaDeviceSynchronize();
        for (auto& b : kv.second) cudaFree(b.second);
        kv.second.clear();
    }
    std::lock_guard<std::mutex> lk2(g_hstage_mu);
    for (auto& b : g_hstage_free) cudaFreeHost(b.first);
    g_hstage_free.clear();
    return PK_OK;
}

template <typename T>
static int dev_alloc(T** p, size_t count) {
    return pool_alloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
}

template <typename T>
static void dev_free(T*& p) {
    pool_free((void*)p);
    p = nullptr;
}

// ---------------------------------------------------------------------------
// Poisson table: per-device copy of the host table
// ---------------------------------------------------------------------------
namespace {
struct DevTable { double* d = nullptr; int32_t kmax = -1; };
std::mutex g_dev_tab_mu;
std::map<int, DevTable> g_dev_tab;
}  // namespace

int pk_poisson_table_device(int device, int32_t k_min_size, const double** d_out, int32_t* k_max_out) {
    std::lock_guard<std::mutex> lk(g_dev_tab_mu);
    DevTable& t = g_dev_tab[device];
    if (t.kmax < k_min_size) {
        int32_t want = std::max(k_min_size, 8191);
        if (t.kmax >= 0) want = std::max(want, 2 * t.kmax + 1);
        const double* host = nullptr;
        PK_CHECK(pk_poisson_table_host(want, &host));
        PK_CUDA(cudaSetDevice(device));
        PK_CUDA(cudaDeviceSynchronize());      // nobody may still read the old copy
        dev_free(t.d);
        PK_CHECK(dev_alloc(&t.d, (size_t)want + 1));
        PK_CUDA(cudaMemcpy(t.d, host, ((size_t)want + 1) * sizeof(double), cudaMemcpyHostToDevice));
        t.kmax = want;
    }
    *d_out = t.d;
    *k_max_out = t.kmax;
    return PK_OK;
}

// ---------------------------------------------------------------------------
// forest
// ---------------------------------------------------------------------------
static float round_down_f32(double t) {
    float f = (float)t;
    if ((double)f > t) f = std::nextafterf(f, -INFINITY);
    return f;
}

extern "C" int pk_forest_create(int device, int32_t n_trees, int32_t n_features, const int64_t* node_offset,
                                const int32_t* feature, const double* threshold, const int32_t* left,
                                const int32_t* right, const uint8_t* missing_left, const double* leaf_p1,
                                pk_forest** out) {
    if (!out || !node_offset || !feature || !threshold || !left || !right || !leaf_p1) {
        pk_set_error("pk_forest_create: NULL argument");
        return PK_EINVAL;
    }
    if (n_trees <= 0 || n_features <= 0 || n_features >= (1 << PK_FEAT_BITS)) {
        pk_set_error("pk_forest_create: n_trees=%d n_features=%d unsupported (features < %d)", n_trees, n_features,
                     1 << PK_FEAT_BITS);
        return PK_EINVAL;
    }
    const int64_t total = node_offset[n_trees];
    if (total <= 0 || total >= (1LL << 31)) { pk_set_error("pk_forest_create: bad node count"); return PK_EINVAL; }
    std::vector<uint2> nodes((size_t)total);
    std::vector<int32_t> orig((size_t)total);
    std::vector<uint32_t> roots((size_t)n_trees);
    std::vector<uint8_t> tdepth((size_t)n_trees);
    std::vector<uint2> nodes_f0((size_t)total), nodes_f1((size_t)total);     // fused-kernel encodings (pk_common.cuh)
    std::vector<uint8_t> rootfeat((size_t)n_trees);
    bool fused_ok = true, cf_ok = n_features <= 256;
    int32_t max_depth = 0;
    std::vector<int32_t> newid, stack, depth;
    for (int32_t t = 0; t < n_trees; ++t) {
        const int64_t o = node_offset[t], cnt = node_offset[t + 1] - o;
        if (cnt <= 0) { pk_set_error("pk_forest_create: tree %d is empty", t); return PK_EINVAL; }
        // preorder renumbering: left child directly follows its parent
        newid.assign((size_t)cnt, -1);
        stack.clear(); depth.assign((size_t)cnt, 0);
        stack.push_back(0);
        int32_t next = 0, this_depth = 0;
        while (!stack.empty()) {
            int32_t v = stack.back(); stack.pop_back();
            if (v < 0 || v >= cnt || newid[v] != -1) { pk_set_error("pk_forest_create: tree %d is not a tree", t); return PK_EINVAL; }
            newid[v] = next++;
            int32_t l = left[o + v], r = right[o + v];
            if (l != -1) {
                if (r < 0 || r >= cnt || l < 0 || l >= cnt) { pk_set_error("pk_forest_create: bad child index"); return PK_EINVAL; }
                depth[l] = depth[r] = depth[v] + 1;
                max_depth = std::max(max_depth, depth[v] + 1);
                this_depth = std::max(this_depth, depth[v] + 1);
                stack.push_back(r);
                stack.push_back(l);
            }
        }
        if (next != cnt) { pk_set_error("pk_forest_create: tree %d has unreachable nodes", t); return PK_EINVAL; }
        for (int32_t v = 0; v < cnt; ++v) {
            const int64_t p = o + newid[v];
            orig[(size_t)p] = v;
            int32_t l = left[o + v], r = right[o + v];
            if (l == -1) {
                double val = leaf_p1[o + v];
                if (!(val >= 0.0) || std::signbit(val)) { pk_set_error("pk_forest_create: leaf value %g is not a fraction", val); return PK_EINVAL; }
                memcpy(&nodes[(size_t)p], &val, 8);
            } else {
                int32_t ft = feature[o + v];
                if (ft < 0 || ft >= n_features) { pk_set_error("pk_forest_create: feature index %d out of range", ft); return PK_EINVAL; }
                if (newid[l] != newid[v] + 1) { pk_set_error("pk_forest_create: internal preorder error"); return PK_EINVAL; }
                uint32_t roff = (uint32_t)(newid[r] - newid[v]);
                if (roff >= (1u << 18)) { pk_set_error("pk_forest_create: tree %d too large (right offset %u)", t, roff); return PK_EUNSUPPORTED; }
                float thr = round_down_f32(threshold[o + v]);
                uint32_t meta = 0x80000000u | ((missing_left && missing_left[o + v]) ? (1u << 30) : 0u) |
                                (roff << 12) | ((uint32_t)ft << 2);
                uint32_t tb;
                memcpy(&tb, &thr, 4);
                nodes[(size_t)p] = make_uint2(tb, meta);
                const uint32_t fl = left[o + l] != -1 ? (uint32_t)feature[o + l] : 0u;
                const uint32_t fr = left[o + r] != -1 ? (uint32_t)feature[o + r] : 0u;
                if (roff >= 4096u) fused_ok = false;
                const uint32_t top = 0x80000000u | ((roff & 4095u) << 19);
                nodes_f0[(size_t)p] = make_uint2(tb, top | ((meta >> 30) & 1u) << 15 | ((uint32_t)ft << 2));
                nodes_f1[(size_t)p] = make_uint2(tb, top | ((fr & 255u) << 8) | (fl & 255u));
            }
        }
        for (int32_t v = 0; v < cnt; ++v)
            if (left[o + v] == -1) nodes_f0[(size_t)(o + newid[v])] = nodes_f1[(size_t)(o + newid[v])] = nodes[(size_t)(o + newid[v])];
        rootfeat[(size_t)t] = left[o] != -1 ? (uint8_t)(feature[o] & 255) : 0;
        roots[(size_t)t] = (uint32_t)o;
        if (this_depth > 255) { pk_set_error("pk_forest_create: tree %d deeper than 255", t); return PK_EUNSUPPORTED; }
        tdepth[(size_t)t] = (uint8_t)this_depth;
    }
    PK_CUDA(cudaSetDevice(device));
    pk_forest* f = new pk_forest();
    f->device = device; f->n_trees = n_trees; f->n_features = n_features; f->n_nodes = total; f->max_depth = max_depth;
    int r;
    f->h_node_offset.assign(node_offset, node_offset + n_trees + 1);
    if ((r = dev_alloc(&f->d_nodes, (size_t)total + 4)) || (r = dev_alloc(&f->d_root, (size_t)n_trees)) ||
        (r = dev_alloc(&f->d_orig, (size_t)total)) || (r = dev_alloc(&f->d_depth, (size_t)n_trees))) {
        pk_forest_destroy(f);
        return r;
    }
    PK_CUDA(cudaMemcpy(f->d_nodes, nodes.data(), (size_t)total * sizeof(uint2), cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemcpy(f->d_root, roots.data(), (size_t)n_trees * sizeof(uint32_t), cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemcpy(f->d_orig, orig.data(), (size_t)total * sizeof(int32_t), cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemset(f->d_nodes + total, 0, 4 * sizeof(uint2)));
    PK_CUDA(cudaMemcpy(f->d_depth, tdepth.data(), (size_t)n_trees, cudaMemcpyHostToDevice));
    f->fused_ok = fused_ok;
    f->cf_ok = fused_ok && cf_ok;
    if (fused_ok) {
        if ((r = dev_alloc(&f->d_nodes_f0, (size_t)total + 4)) || (r = dev_alloc(&f->d_nodes_f1, (size_t)total + 4)) ||
            (r = dev_alloc(&f->d_rootfeat, (size_t)n_trees))) {
            pk_forest_destroy(f);
            return r;
        }
        PK_CUDA(cudaMemcpy(f->d_nodes_f0, nodes_f0.data(), (size_t)total * sizeof(uint2), cudaMemcpyHostToDevice));
        PK_CUDA(cudaMemset(f->d_nodes_f0 + total, 0, 4 * sizeof(uint2)));
        PK_CUDA(cudaMemcpy(f->d_nodes_f1, nodes_f1.data(), (size_t)total * sizeof(uint2), cudaMemcpyHostToDevice));
        PK_CUDA(cudaMemset(f->d_nodes_f1 + total, 0, 4 * sizeof(uint2)));
        PK_CUDA(cudaMemcpy(f->d_rootfeat, rootfeat.data(), (size_t)n_trees, cudaMemcpyHostToDevice));
    }
    *out = f;
    return PK_OK;
}

int pk_forest_groups(pk_forest* f, int tbn, int chunk, const int4** d_groups, int32_t* n_groups) {
    for (auto& g : f->group_tables)
        if (g.tbn == tbn && g.chunk == chunk) { *d_groups = g.d; *n_groups = g.n; return PK_OK; }
    const std::vector<int64_t>& off = f->h_node_offset;
    const int32_t n_trees = f->n_trees;
    std::vector<int4> groups;
    for (int32_t t = 0; t < n_trees;) {
        int64_t gbase = off[t] & ~1LL;
        int32_t t1 = t + 1;
        while (t1 < n_trees && off[t1 + 1] - gbase <= tbn) ++t1;
        if (t1 < n_trees && t1 - t > chunk && (t1 - t) % chunk) t1 -= (t1 - t) % chunk;
        const bool fits = off[t1] - gbase <= tbn;
        int64_t staged = std::min<int64_t>(off[t1] - gbase, tbn);
        staged = (staged + 1) & ~1LL;                       // 16-byte multiple; the node array is padded
        groups.push_back(make_int4(t, t1 - t, (int)gbase, fits ? (int)staged : -(int)staged));
        t = t1;
    }
    pk_forest::GroupTable gt;
    gt.tbn = tbn; gt.chunk = chunk; gt.n = (int32_t)groups.size();
    PK_CUDA(cudaSetDevice(f->device));
    PK_CHECK(dev_alloc(&gt.d, groups.size()));
    PK_CUDA(cudaMemcpy(gt.d, groups.data(), groups.size() * sizeof(int4), cudaMemcpyHostToDevice));
    f->group_tables.push_back(gt);
    *d_groups = gt.d; *n_groups = gt.n;
    return PK_OK;
}

extern "C" int pk_forest_destroy(pk_forest* f) {
    if (!f) return PK_OK;
    cudaSetDevice(f->device);
    dev_free(f->d_nodes); dev_free(f->d_root); dev_free(f->d_orig); dev_free(f->d_depth);
    dev_free(f->d_nodes_f0); dev_free(f->d_nodes_f1); dev_free(f->d_rootfeat);
    for (auto& g : f->group_tables) dev_free(g.d);
    delete f;
    return PK_OK;
}

extern "C" int pk_forest_info(const pk_forest* f, int32_t* n_trees, int32_t* n_features, int64_t* n_nodes) {
    if (!f) { pk_set_error("pk_forest_info: NULL forest"); return PK_EINVAL; }
    if (n_trees) *n_trees = f->n_trees;
    if (n_features) *n_features = f->n_features;
    if (n_nodes) *n_nodes = f->n_nodes;
    return PK_OK;
}

extern "C" int pk_forest_apply(pk_forest* f, const float* X, int64_t n_rows, int32_t* leaves, double* proba, void* stream) {
    if (!f || !X || n_rows < 0) { pk_set_error("pk_forest_apply: bad argument"); return PK_EINVAL; }
    PK_CUDA(cudaSetDevice(f->device));
    return pk_launch_forest(f, X, nullptr, n_rows, leaves, proba, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// chromosome
// ---------------------------------------------------------------------------
static int64_t band_pixels_of(const pk_chrom* c) {
    int64_t k = (int64_t)c->upper - c->lower + 1;
    if (k <= 0) return 0;
    return k * c->n - ((int64_t)c->lower + c->upper) * k / 2;
}

// Skewed tensor map over the row-major band copy (pk_common.cuh): element (j, i) = band2[j * (P2 - 1) + i], the
// dense matrix cell (row j, column i). The driver entry point is looked up at run time (no link against libcuda).
typedef CUresult (*pk_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_band2_map(pk_chrom* c) {
    static pk_encode_tiled_fn enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            pk_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return PK_EUNSUPPORTED;
        }
        enc = (pk_encode_tiled_fn)fn;
    }
    const int S = c->S, BC = (S + 3 + 3) & ~3;        // box columns: the window plus the slack of a 16-byte aligned start
    const cuuint64_t dims[2] = {(cuuint64_t)c->n + 8, (cuuint64_t)c->n};
    const cuuint64_t strides[1] = {(cuuint64_t)(c->P2 - 1) * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BC, (cuuint32_t)S}, estr[2] = {1, 1};
    const CUresult r = enc(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, c->d_band2, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { pk_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return PK_ECUDA; }
    c->tmap_ok = true;
    return PK_OK;
}

extern "C" int pk_chrom_create(int device, int32_t n_bins, int32_t width, int32_t lower, int32_t upper, int balanced,
                               void* stream, pk_chrom** out) {
    if (!out) { pk_set_error("pk_chrom_create: out is NULL"); return PK_EINVAL; }
    if (n_bins <= 0 || width < 1 || width > PK_MAX_W) {
        pk_set_error("pk_chrom_create: n_bins=%d width=%d unsupported (1 <= width <= %d)", n_bins, width, PK_MAX_W);
        return PK_EINVAL;
    }
    int32_t lo = std::max(lower, width + 1);               // scoreUtils.py:13
    int32_t up = std::min(upper, n_bins - 2 * width);      // scoreUtils.py:14
    if (up + 2 * width < 0) {
        pk_set_error("pk_chrom_create: chromosome of %d bins is too small for width %d", n_bins, width);
        return PK_EINVAL;
    }
    PK_CUDA(cudaSetDevice(device));
    pk_chrom* c = new pk_chrom();
    c->device = device; c->stream = (cudaStream_t)stream;
    c->n = n_bins; c->w = width; c->S = 2 * width + 1; c->F = c->S * c->S;
    c->lower = lo; c->upper = up; c->ND = up + 2 * width + 1;
    c->pitch = ((int64_t)n_bins + 31) / 32 * 32;
    c->balanced = balanced ? 1 : 0;
    c->row_begin = 0; c->row_end = n_bins;
    int r = PK_OK;
    const size_t bandsz = (size_t)c->ND * (size_t)c->pitch;
    if ((r = dev_alloc(&c->d_band, bandsz)) || (r = dev_alloc(&c->d_w, (size_t)n_bins)) ||
        (r = dev_alloc(&c->d_valid, (size_t)n_bins)) || (r = dev_alloc(&c->d_vbits, (size_t)n_bins / 32 + 4)) ||
        (r = dev_alloc(&c->d_scratch, bandsz)) ||
        (r = dev_alloc(&c->d_diag_sum, (size_t)c->ND)) || (r = dev_alloc(&c->d_diag_cnt, (size_t)c->ND)) ||
        (r = dev_alloc(&c->d_exp, (size_t)c->ND)) || (r = dev_alloc(&c->d_bg, (size_t)c->ND)) ||
        (r = dev_alloc(&c->d_rowptr, (size_t)n_bins + 1))) {
        pk_chrom_destroy(c);
        return r;
    }
    // Everything the host reads back after a scoring pass sits in one block, so that one copy fetches it:
    // flags int32[4] | candidate totals int64[2] | counters uint64[4] | surviving windows per reference batch int32[nb]
    // (batches are numbered over the whole chromosome's candidates: at most the band pixels)
    c->batch_cap = band_pixels_of(c) / PK_BATCH + 2;
    c->head_bytes = 64 + 4 * (size_t)c->batch_cap;
    if ((r = dev_alloc(&c->d_head, c->head_bytes))) { pk_chrom_destroy(c); return r; }
    c->d_flags = reinterpret_cast<int32_t*>(c->d_head);
    c->d_ncand = reinterpret_cast<long long*>(c->d_head + 16);
    c->d_counters = reinterpret_cast<unsigned long long*>(c->d_head + 32);
    c->d_batch_win = reinterpret_cast<int32_t*>(c->d_head + 64);
    for (auto& e : c->ev) {
        if (cudaEventCreate(&e) != cudaSuccess) { pk_set_error("cudaEventCreate failed"); pk_chrom_destroy(c); return PK_ECUDA; }
    }
    if ((g_tune_tma == 1 && width == 7) || (g_tune_tma > 1 && (width == 5 || width == 7))) {
        // row-major band copy + tensor map for the fused kernel's window fetch; without them (an old driver)
        // the kernel gathers cell by cell
        c->P2 = ((int64_t)c->ND + 3) / 4 * 4 + 1;
        if (dev_alloc(&c->d_band2, (size_t)c->n * (size_t)c->P2 + 64) == PK_OK) {
            if (make_band2_map(c) != PK_OK) dev_free(c->d_band2);
        } else {
            cudaGetLastError();
        }
    }
    *out = c;
    return PK_OK;
}

extern "C" int pk_chrom_destroy(pk_chrom* c) {
    if (!c) return PK_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream); else cudaDeviceSynchronize();
    dev_free(c->d_band); dev_free(c->d_band2); dev_free(c->d_w); dev_free(c->d_wp); dev_free(c->d_valid); dev_free(c->d_vbits); dev_free(c->d_scratch);
    dev_free(c->d_diag_sum); dev_free(c->d_diag_cnt);
    dev_free(c->d_exp); dev_free(c->d_bg); dev_free(c->d_head);
    c->d_flags = nullptr; c->d_counters = nullptr; c->d_ncand = nullptr; c->d_batch_win = nullptr;
    dev_free(c->d_rowptr);
    dev_free(c->d_b1); dev_free(c->d_b2); dev_free(c->d_cnt); dev_free(c->d_blob_own);
    dev_free(c->d_cstate); dev_free(c->d_bits);
    dev_free(c->d_cx); dev_free(c->d_cd); dev_free(c->d_crank);
    dev_free(c->d_keep); dev_free(c->d_fea32); dev_free(c->d_prob);
    dev_free(c->d_rx); dev_free(c->d_ry); dev_free(c->d_rb); dev_free(c->d_rp); dev_free(c->d_rv);
    if (c->ev_x) cudaEventDestroy(c->ev_x);
    dev_free(c->d_rowcnt); dev_free(c->d_rowoff); dev_free(c->d_rrank); dev_free(c->d_perm); dev_free(c->d_packed);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    hstage_release(c->h_stage, c->h_stage_bytes);
    delete c;
    return PK_OK;
}

extern "C" int pk_chrom_bounds(const pk_chrom* c, int32_t* lower_eff, int32_t* upper_eff, int32_t* exp_len) {
    if (!c) { pk_set_error("pk_chrom_bounds: NULL handle"); return PK_EINVAL; }
    if (lower_eff) *lower_eff = c->lower;
    if (upper_eff) *upper_eff = c->upper;
    if (exp_len) *exp_len = c->ND;
    return PK_OK;
}

static float ev_ms(pk_chrom* c, int a, int b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.f; }
    return ms;
}

static int stage_pixels(pk_chrom* c, const int32_t** p, int32_t** d_stage, int64_t nnz, int mem) {
    // host column -> staging buffer on the device (grow-only)
    if (mem != PK_MEM_HOST || nnz == 0) return PK_OK;
    PK_CUDA(cudaMemcpyAsync(*d_stage, *p, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
    *p = *d_stage;
    return PK_OK;
}

// Wait for everything queued on the handle's streams: called before a buffer that queued work may
// still read is returned to the block cache (growth of a reused handle's buffers; rare).
static int quiesce(pk_chrom* c) {
    PK_CUDA(cudaStreamSynchronize(c->stream));
    if (c->score_stream) PK_CUDA(cudaStreamSynchronize(c->score_stream));
    return PK_OK;
}

static int reserve_staging(pk_chrom* c, int64_t nnz, bool need_b1) {
    if (nnz > c->pix_cap || !c->d_b2 || (need_b1 && !c->d_b1)) {
        if (c->d_b2) PK_CHECK(quiesce(c));
        dev_free(c->d_b1); dev_free(c->d_b2); dev_free(c->d_cnt);
        int64_t cap = std::max<int64_t>(nnz, 1024);
        PK_CHECK(dev_alloc(&c->d_b1, (size_t)cap)); PK_CHECK(dev_alloc(&c->d_b2, (size_t)cap)); PK_CHECK(dev_alloc(&c->d_cnt, (size_t)cap));
        c->pix_cap = cap;
    }
    return PK_OK;
}

static int after_band(pk_chrom* c) {
    cudaStream_t s = c->stream;
    c->band2_valid = false;
    if (c->d_band2 && c->tmap_ok && g_tune_tma) PK_CHECK(pk_launch_band_rowmajor(c));     // while the band is hot in L2
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[1], s));
    PK_CHECK(pk_launch_diag_sums(c));
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[2], s));
    c->has_pixels = true; c->has_expected = false; c->has_candidates = false; c->has_scores = false;
    c->use_wp = false;
    return PK_OK;
}

// Weights of the Poisson filter alone (scoreUtils.py:55-57) when they are not the ones that balance the
// pixel values: cooler inverts a weight column flagged `divisive_weights` inside matrix(balance=...), while
// the reference hands the column's raw values to Chromosome (score_chromosome.py:44).
extern "C" int pk_chrom_set_poisson_weights(pk_chrom* c, const double* weights, int mem) {
    if (!c || !weights) { pk_set_error("pk_chrom_set_poisson_weights: bad argument"); return PK_EINVAL; }
    if (!c->balanced) { pk_set_error("pk_chrom_set_poisson_weights: raw mode has no weights"); return PK_ESTATE; }
    if (!c->has_pixels) { pk_set_error("pk_chrom_set_poisson_weights: upload the pixels first"); return PK_ESTATE; }
    PK_CUDA(cudaSetDevice(c->device));
    if (!c->d_wp) PK_CHECK(dev_alloc(&c->d_wp, (size_t)c->n));
    PK_CUDA(cudaMemcpyAsync(c->d_wp, weights, (size_t)c->n * sizeof(double),
                            (mem & ~PK_PIXELS_SORTED) == PK_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
    c->use_wp = true; c->has_candidates = false; c->has_scores = false;
    return PK_OK;
}

extern "C" int pk_chrom_upload_pixels(pk_chrom* c, const int32_t* bin1, const int32_t* bin2, const int32_t* count,
                                      int64_t nnz, const double* weights, int mem) {
    if (!c || nnz < 0 || (nnz > 0 && (!bin1 || !bin2 || !count))) { pk_set_error("pk_chrom_upload_pixels: bad argument"); return PK_EINVAL; }
    if (c->balanced && !weights) { pk_set_error("pk_chrom_upload_pixels: balanced mode needs weights"); return PK_EINVAL; }
    const bool sorted = (mem & PK_PIXELS_SORTED) != 0;
    mem &= ~PK_PIXELS_SORTED;
    PK_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int32_t *p1 = bin1, *p2 = bin2, *pc = count;
    if (mem == PK_MEM_HOST) {
        PK_CHECK(reserve_staging(c, nnz, true));
        PK_CHECK(stage_pixels(c, &p1, &c->d_b1, nnz, mem));
        PK_CHECK(stage_pixels(c, &p2, &c->d_b2, nnz, mem));
        PK_CHECK(stage_pixels(c, &pc, &c->d_cnt, nnz, mem));
    }
    if (c->balanced)
        PK_CUDA(cudaMemcpyAsync(c->d_w, weights, (size_t)c->n * sizeof(double),
                                mem == PK_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[0], s));
    PK_CUDA(cudaMemsetAsync(c->d_valid, 0, (size_t)c->n, s));
    PK_CUDA(cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int32_t), s));
    c->declared_sorted = sorted;
    c->up_kind = sorted ? 2 : 1; c->up_b1 = p1; c->up_b2 = p2; c->up_cnt = pc; c->up_rowptr = c->d_rowptr; c->up_nnz = nnz;
    if (sorted) {
        // cooler order: derive row offsets on the device, then the tiled CSR build
        PK_CHECK(pk_launch_rowptr(c, p1, p2, nnz, c->d_rowptr));
        PK_CHECK(pk_launch_band_csr(c, c->d_rowptr, p2, pc, 0));
    } else {
        PK_CUDA(cudaMemsetAsync(c->d_band, 0, (size_t)c->ND * c->pitch * sizeof(int32_t), s));
        PK_CHECK(pk_launch_scatter(c, p1, p2, pc, nnz));
    }
    return after_band(c);
}

extern "C" int pk_chrom_upload_csr(pk_chrom* c, const int64_t* bin1_offset, const int32_t* bin2, const int32_t* count,
                                   int64_t nnz, const double* weights, int mem) {
    if (!c || nnz < 0 || !bin1_offset || (nnz > 0 && (!bin2 || !count))) { pk_set_error("pk_chrom_upload_csr: bad argument"); return PK_EINVAL; }
    if (c->balanced && !weights) { pk_set_error("pk_chrom_upload_csr: balanced mode needs weights"); return PK_EINVAL; }
    mem &= ~PK_PIXELS_SORTED;
    PK_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int32_t *p2 = bin2, *pc = count;
    const long long* rp = reinterpret_cast<const long long*>(bin1_offset);
    if (mem == PK_MEM_HOST) {
        PK_CHECK(reserve_staging(c, nnz, false));
        PK_CHECK(stage_pixels(c, &p2, &c->d_b2, nnz, mem));
        PK_CHECK(stage_pixels(c, &pc, &c->d_cnt, nnz, mem));
        PK_CUDA(cudaMemcpyAsync(c->d_rowptr, bin1_offset, ((size_t)c->n + 1) * 8, cudaMemcpyHostToDevice, s));
        rp = c->d_rowptr;
    }
    if (c->balanced)
        PK_CUDA(cudaMemcpyAsync(c->d_w, weights, (size_t)c->n * sizeof(double),
                                mem == PK_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[0], s));
    PK_CUDA(cudaMemsetAsync(c->d_valid, 0, (size_t)c->n, s));
    PK_CUDA(cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int32_t), s));
    c->declared_sorted = false;      // row offsets given: nothing to verify
    c->up_kind = 2; c->up_b1 = nullptr; c->up_b2 = p2; c->up_cnt = pc; c->up_rowptr = rp; c->up_nnz = nnz;
    PK_CHECK(pk_launch_band_csr(c, rp, p2, pc, 0));
    return after_band(c);
}

extern "C" int pk_chrom_upload_csr16(pk_chrom* c, const int64_t* bin1_offset, const uint16_t* bin2_delta, const uint16_t* count,
                                     int64_t nnz, const double* weights, int mem) {
    if (!c || nnz < 0 || !bin1_offset || (nnz > 0 && (!bin2_delta || !count))) { pk_set_error("pk_chrom_upload_csr16: bad argument"); return PK_EINVAL; }
    if (c->balanced && !weights) { pk_set_error("pk_chrom_upload_csr16: balanced mode needs weights"); return PK_EINVAL; }
    mem &= ~PK_PIXELS_SORTED;
    PK_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const void *p2 = bin2_delta, *pc = count;
    const long long* rp = reinterpret_cast<const long long*>(bin1_offset);
    if (mem == PK_MEM_HOST) {
        PK_CHECK(reserve_staging(c, (nnz + 1) / 2, false));        // two 16-bit values per staged int32
        if (nnz > 0) {
            PK_CUDA(cudaMemcpyAsync(c->d_b2, bin2_delta, (size_t)nnz * 2, cudaMemcpyHostToDevice, s));
            PK_CUDA(cudaMemcpyAsync(c->d_cnt, count, (size_t)nnz * 2, cudaMemcpyHostToDevice, s));
        }
        p2 = c->d_b2; pc = c->d_cnt;
        PK_CUDA(cudaMemcpyAsync(c->d_rowptr, bin1_offset, ((size_t)c->n + 1) * 8, cudaMemcpyHostToDevice, s));
        rp = c->d_rowptr;
    }
    if (c->balanced)
        PK_CUDA(cudaMemcpyAsync(c->d_w, weights, (size_t)c->n * sizeof(double),
                                mem == PK_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[0], s));
    PK_CUDA(cudaMemsetAsync(c->d_valid, 0, (size_t)c->n, s));
    PK_CUDA(cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int32_t), s));
    c->declared_sorted = false;
    c->up_kind = 3; c->up_b1 = nullptr; c->up_b2 = p2; c->up_cnt = pc; c->up_rowptr = rp; c->up_nnz = nnz;
    PK_CHECK(pk_launch_band_csr(c, rp, p2, pc, 1));
    return after_band(c);
}

// Packed pixel rows (one contiguous blob, layout in the header file): a chromosome crosses the bus in one copy.
extern "C" int pk_chrom_upload_rows(pk_chrom* c, const void* blob, int64_t bytes, const double* weights, int mem) {
    if (!c || !blob || bytes < 128) { pk_set_error("pk_chrom_upload_rows: bad argument"); return PK_EINVAL; }
    if (c->balanced && !weights) { pk_set_error("pk_chrom_upload_rows: balanced mode needs weights"); return PK_EINVAL; }
    mem &= ~PK_PIXELS_SORTED;
    PK_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    long long* h = c->rows_hdr;
    if (mem == PK_MEM_HOST) memcpy(h, blob, 128);
    else PK_CUDA(cudaMemcpy(h, blob, 128, cudaMemcpyDeviceToHost));
    if (h[0] != PK_ROWS_MAGIC) { pk_set_error("pk_chrom_upload_rows: not a packed-rows blob"); return PK_EINVAL; }
    if (h[1] != c->n) { pk_set_error("pk_chrom_upload_rows: blob holds %lld bins, the handle %d", h[1], c->n); return PK_EINVAL; }
    if (h[2] < c->ND) {
        pk_set_error("pk_chrom_upload_rows: blob packs %lld distances, the band needs upper + 2w + 1 = %d", h[2], c->ND);
        return PK_EINVAL;
    }
    if (h[14] > bytes || h[3] != (h[2] + 31) / 32 || h[4] < 0 || h[5] < 0 || h[6] < 0) {
        pk_set_error("pk_chrom_upload_rows: inconsistent header"); return PK_EINVAL;
    }
    const long long n1 = (long long)c->n + 1;
    const long long need[7] = {h[7] + 4 * h[3] * c->n, h[8] + 4 * n1, h[9] + h[4], h[10] + 12 * h[5], h[11] + 8 * n1,
                               h[12] + 4 * h[6], h[13] + 4 * h[6]};
    for (int i = 0; i < 7; ++i)
        if (h[7 + i] < 128 || (h[7 + i] & 15) || need[i] > h[14]) { pk_set_error("pk_chrom_upload_rows: section %d out of bounds", i); return PK_EINVAL; }
    if (mem == PK_MEM_HOST) {
        if (h[14] > c->blob_cap || !c->d_blob_own) {
            if (c->d_blob_own) PK_CHECK(quiesce(c));
            dev_free(c->d_blob_own);
            PK_CHECK(dev_alloc(&c->d_blob_own, (size_t)h[14] + (size_t)h[14] / 8));
            c->blob_cap = h[14] + h[14] / 8;
        }
        PK_CUDA(cudaMemcpyAsync(c->d_blob_own, blob, (size_t)h[14], cudaMemcpyHostToDevice, s));
        c->d_blob = c->d_blob_own;
    } else {
        c->d_blob = (unsigned char*)blob;
    }
    if (c->balanced)
        PK_CUDA(cudaMemcpyAsync(c->d_w, weights, (size_t)c->n * sizeof(double),
                                mem == PK_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[0], s));
    PK_CUDA(cudaMemsetAsync(c->d_valid, 0, (size_t)c->n, s));
    PK_CUDA(cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int32_t), s));
    c->declared_sorted = false;
    c->up_kind = 4; c->up_b1 = nullptr; c->up_b2 = nullptr; c->up_cnt = nullptr; c->up_rowptr = nullptr; c->up_nnz = h[4] + h[6];
    PK_CHECK(pk_launch_band_rows(c));
    return after_band(c);
}

// calculate_depth.py:25-28: sum of the raw counts of the uploaded pixels with bin2 - bin1 >= min_dis_bins
extern "C" int pk_chrom_depth(pk_chrom* c, int32_t min_dis_bins, int64_t* total) {
    if (!c || !total) { pk_set_error("pk_chrom_depth: bad argument"); return PK_EINVAL; }
    if (!c->has_pixels || c->up_kind == 0) { pk_set_error("pk_chrom_depth: no pixels uploaded"); return PK_ESTATE; }
    PK_CUDA(cudaSetDevice(c->device));
    unsigned long long* d_total = c->d_counters + 3;          // spare counter slot
    PK_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), c->stream));
    PK_CHECK(pk_launch_depth(c, min_dis_bins, d_total));
    unsigned long long h = 0;
    PK_CUDA(cudaMemcpyAsync(&h, d_total, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaStreamSynchronize(c->stream));
    *total = (int64_t)h;
    return PK_OK;
}

extern "C" int pk_chrom_diag_sums(pk_chrom* c, double* out_sum, int64_t* out_cnt) {
    if (!c || !out_sum || !out_cnt) { pk_set_error("pk_chrom_diag_sums: bad argument"); return PK_EINVAL; }
    if (!c->has_pixels) { pk_set_error("pk_chrom_diag_sums: no pixels uploaded"); return PK_ESTATE; }
    PK_CUDA(cudaSetDevice(c->device));
    PK_CUDA(cudaMemcpyAsync(out_sum, c->d_diag_sum, (size_t)c->ND * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaMemcpyAsync(out_cnt, c->d_diag_cnt, (size_t)c->ND * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaStreamSynchronize(c->stream));
    return PK_OK;
}

extern "C" int pk_chrom_set_expected(pk_chrom* c, const double* exp_arr, const double* background) {
    if (!c || !exp_arr || !background) { pk_set_error("pk_chrom_set_expected: bad argument"); return PK_EINVAL; }
    PK_CUDA(cudaSetDevice(c->device));
    PK_CUDA(cudaMemcpyAsync(c->d_exp, exp_arr, (size_t)c->ND * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PK_CUDA(cudaMemcpyAsync(c->d_bg, background, (size_t)c->ND * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PK_CUDA(cudaStreamSynchronize(c->stream));   // the host arrays may be pageable temporaries
    c->has_expected = true; c->has_candidates = false; c->has_scores = false;
    return PK_OK;
}

extern "C" int pk_chrom_fit_expected(pk_chrom* c) {
    if (!c) { pk_set_error("pk_chrom_fit_expected: NULL handle"); return PK_EINVAL; }
    if (!c->has_pixels) { pk_set_error("pk_chrom_fit_expected: no pixels uploaded"); return PK_ESTATE; }
    PK_CUDA(cudaSetDevice(c->device));
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[3], c->stream));
    if (pk_fit_on_device_supported(c->ND)) {
        PK_CHECK(pk_launch_fit_expected(c));      // stays on the stream; a failed fit raises flags[2]
    } else {
        std::vector<double> sum((size_t)c->ND), e((size_t)c->ND);
        std::vector<long long> cnt((size_t)c->ND);
        PK_CUDA(cudaMemcpyAsync(sum.data(), c->d_diag_sum, (size_t)c->ND * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        PK_CUDA(cudaMemcpyAsync(cnt.data(), c->d_diag_cnt, (size_t)c->ND * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        PK_CUDA(cudaStreamSynchronize(c->stream));
        PK_CHECK(pk_fit_expected_host(sum.data(), cnt.data(), c->ND, e.data()));
        PK_CUDA(cudaMemcpyAsync(c->d_exp, e.data(), (size_t)c->ND * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        PK_CUDA(cudaMemcpyAsync(c->d_bg, e.data(), (size_t)c->ND * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        PK_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[4], c->stream));
    c->has_expected = true; c->has_candidates = false; c->has_scores = false;
    return PK_OK;
}

// device flags -> host; turns raised flags into errors where nothing can be retried
static int read_flags(pk_chrom* c, int32_t flags[4]) {
    PK_CUDA(cudaMemcpyAsync(flags, c->d_flags, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaStreamSynchronize(c->stream));
    if (c->declared_sorted && (flags[3] & 1)) {
        pk_set_error("pixels were passed with PK_PIXELS_SORTED but are not sorted by (bin1, bin2) with bin1 <= bin2");
        return PK_EINVAL;
    }
    if (flags[2] & 2) {
        pk_set_error("chromosome too long for the per-diagonal sum tables (more than ~230k bins)");
        return PK_EUNSUPPORTED;
    }
    if (flags[2] & 1) {
        pk_set_error("expected curve: no distance has a positive mean (reference raises in IsotonicRegression.fit)");
        return PK_EINVAL;
    }
    return PK_OK;
}

extern "C" int pk_chrom_get_expected(pk_chrom* c, double* out_exp) {
    if (!c || !out_exp) { pk_set_error("pk_chrom_get_expected: bad argument"); return PK_EINVAL; }
    if (!c->has_expected) { pk_set_error("pk_chrom_get_expected: expected curve not set"); return PK_ESTATE; }
    PK_CUDA(cudaSetDevice(c->device));
    int32_t flags[4];
    PK_CHECK(read_flags(c, flags));
    PK_CUDA(cudaMemcpyAsync(out_exp, c->d_exp, (size_t)c->ND * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaStreamSynchronize(c->stream));
    return PK_OK;
}

static int reserve_candidates(pk_chrom* c, int64_t want) {
    if (want > c->cand_cap || !c->d_cx) {
        if (c->d_cx) PK_CHECK(quiesce(c));
        dev_free(c->d_cx); dev_free(c->d_cd); dev_free(c->d_crank);
        dev_free(c->d_keep); dev_free(c->d_prob);
        dev_free(c->d_rx); dev_free(c->d_ry); dev_free(c->d_rb); dev_free(c->d_rp); dev_free(c->d_rv);
        const int64_t cap = std::max<int64_t>(want, 4096);
        PK_CHECK(dev_alloc(&c->d_cx, (size_t)cap)); PK_CHECK(dev_alloc(&c->d_cd, (size_t)cap)); PK_CHECK(dev_alloc(&c->d_crank, (size_t)cap));
        PK_CHECK(dev_alloc(&c->d_keep, (size_t)cap)); PK_CHECK(dev_alloc(&c->d_prob, (size_t)cap));
        PK_CHECK(dev_alloc(&c->d_rx, (size_t)cap)); PK_CHECK(dev_alloc(&c->d_ry, (size_t)cap)); PK_CHECK(dev_alloc(&c->d_rb, (size_t)cap));
        PK_CHECK(dev_alloc(&c->d_rp, (size_t)cap)); PK_CHECK(dev_alloc(&c->d_rv, (size_t)cap));
        c->cand_cap = cap;
    }
    return PK_OK;
}

// candidate scan on the stream; kmin = smallest Poisson table size wanted
static int run_candidates(pk_chrom* c, int32_t kmin) {
    cudaStream_t s = c->stream;
    const int nd = c->upper - c->lower + 1;
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[5], s));
    PK_CUDA(cudaMemsetAsync(c->d_ncand, 0, 2 * sizeof(long long), s));
    if (nd > 0) {
        c->n_chunks = (c->n + 4095) / 4096;               // PK_CTILE slots per scan tile
        const int64_t m = (int64_t)nd * c->n_chunks;
        if (m >= (1LL << 31)) { pk_set_error("candidate scan: %lld tiles", (long long)m); return PK_EUNSUPPORTED; }
        if (m + 1 > c->cnt_cap) {
            if (c->d_cstate) PK_CHECK(quiesce(c));
            dev_free(c->d_cstate); dev_free(c->d_bits);
            PK_CHECK(dev_alloc(&c->d_cstate, (size_t)m + 1));
            PK_CHECK(dev_alloc(&c->d_bits, (size_t)m * 128));
            c->cnt_cap = m + 1;
        }
        const double* d_crit = nullptr;
        int32_t kmax = 0;
        PK_CHECK(pk_poisson_table_device(c->device, kmin, &d_crit, &kmax));
        PK_CHECK(pk_launch_candidates(c, d_crit, kmax));
    }
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[6], s));
    return PK_OK;
}

// read the candidate totals; if a device-side capacity was exceeded, enlarge and redo the scan
static int settle_candidates(pk_chrom* c) {
    for (int attempt = 0; attempt < 4; ++attempt) {
        int32_t flags[4];
        long long nc[2] = {0, 0};
        PK_CUDA(cudaMemcpyAsync(nc, c->d_ncand, sizeof nc, cudaMemcpyDeviceToHost, c->stream));
        PK_CHECK(read_flags(c, flags));
        c->n_cand = nc[0]; c->n_cand_all = nc[1];
        const bool table_small = flags[0] != 0, cand_small = (flags[3] & 2) != 0 || c->n_cand > c->cand_cap;
        if (!table_small && !cand_small) { c->n_cand_known = true; return PK_OK; }
        if (table_small && flags[1] > (1 << 22)) {
            pk_set_error("raw count %d above the Poisson table limit (2^22)", flags[1]);
            return PK_EUNSUPPORTED;
        }
        if (cand_small) PK_CHECK(reserve_candidates(c, c->n_cand + c->n_cand / 8 + 1024));
        // clear the two capacity flags, keep the rest (largest count, weight-range bit)
        const int32_t keep[4] = {0, flags[1], flags[2], flags[3] & ~2};
        PK_CUDA(cudaMemcpyAsync(c->d_flags, keep, sizeof keep, cudaMemcpyHostToDevice, c->stream));
        PK_CHECK(run_candidates(c, table_small ? flags[1] : 0));
    }
    pk_set_error("candidate scan did not settle");
    return PK_ECUDA;
}

extern "C" int pk_chrom_find_candidates(pk_chrom* c, int32_t row_begin, int32_t row_end, int64_t* n_candidates) {
    if (!c) { pk_set_error("pk_chrom_find_candidates: NULL handle"); return PK_EINVAL; }
    if (!c->has_expected) { pk_set_error("pk_chrom_find_candidates: expected curve not set"); return PK_ESTATE; }
    if (row_begin < 0 || row_end > c->n || row_begin > row_end) { pk_set_error("pk_chrom_find_candidates: bad row range [%d, %d)", row_begin, row_end); return PK_EINVAL; }
    PK_CUDA(cudaSetDevice(c->device));
    c->row_begin = row_begin; c->row_end = row_end;
    c->whole = (row_begin == 0 && row_end == c->n);
    c->n_cand_known = false;
    // room for 1/8 of the band pixels of the row tile (synthetic and real maps keep 3-6 %);
    // a device flag reports the rare overflow and the scan is redone with the exact size
    const int64_t tile_px = band_pixels_of(c) * (int64_t)(row_end - row_begin) / std::max(c->n, 1);
    PK_CHECK(reserve_candidates(c, tile_px / 8 + 4096));
    PK_CHECK(run_candidates(c, 0));
    c->has_candidates = true; c->has_scores = false;
    if (n_candidates) {
        PK_CHECK(settle_candidates(c));
        *n_candidates = c->n_cand;
    }
    return PK_OK;
}

extern "C" int pk_chrom_candidates(pk_chrom* c, int32_t* out_x, int32_t* out_y, int64_t capacity, int64_t* n) {
    if (!c || !n) { pk_set_error("pk_chrom_candidates: bad argument"); return PK_EINVAL; }
    if (!c->has_candidates) { pk_set_error("pk_chrom_candidates: find_candidates not called"); return PK_ESTATE; }
    PK_CUDA(cudaSetDevice(c->device));
    if (!c->n_cand_known) PK_CHECK(settle_candidates(c));
    *n = c->n_cand;
    if (!out_x && !out_y) return PK_OK;
    if (capacity < c->n_cand) { pk_set_error("pk_chrom_candidates: capacity %lld < %lld", (long long)capacity, (long long)c->n_cand); return PK_ECAPACITY; }
    if (c->n_cand == 0) return PK_OK;
    std::vector<int32_t> xs((size_t)c->n_cand), d((size_t)c->n_cand);
    PK_CUDA(cudaMemcpyAsync(xs.data(), c->d_cx, (size_t)c->n_cand * 4, cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaMemcpyAsync(d.data(), c->d_cd, (size_t)c->n_cand * 4, cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < c->n_cand; ++i) {
        if (out_x) out_x[i] = xs[(size_t)i];
        if (out_y) out_y[i] = xs[(size_t)i] + d[(size_t)i];
    }
    return PK_OK;
}

static int ensure_feature_buffer(pk_chrom* c) {
    if (c->n_cand > c->fea_cap || !c->d_fea32) {
        if (c->d_fea32) PK_CHECK(quiesce(c));
        dev_free(c->d_fea32);
        int64_t cap = std::max<int64_t>(c->n_cand + c->n_cand / 8, 1024);
        PK_CHECK(dev_alloc(&c->d_fea32, (size_t)cap * c->F));
        c->fea_cap = cap;
    }
    return PK_OK;
}

static int reset_score_state(pk_chrom* c) {
    // counters and the per-batch window counts are neighbours in the head block: one memset
    PK_CUDA(cudaMemsetAsync(c->d_counters, 0, 32 + (size_t)c->batch_cap * 4, c->stream));
    PK_CUDA(cudaMemsetAsync(c->d_keep, 0, (size_t)c->cand_cap, c->stream));
    // buffers of the record ordering that follows the scoring pass
    c->eager_valid = false;
    const int64_t M = c->cand_cap;              // every record is a candidate: the ordering buffers cover them all
    if (!c->d_rowcnt) {
        PK_CHECK(dev_alloc(&c->d_rowcnt, (size_t)c->n + 1)); PK_CHECK(dev_alloc(&c->d_rowoff, (size_t)c->n + 2));
        PK_CUDA(cudaMemsetAsync(c->d_rowcnt, 0, ((size_t)c->n + 1) * 4, c->stream));     // k_record_place keeps it zero afterwards
    }
    if (c->rrank_cap != c->cand_cap || !c->d_rrank) {
        if (c->d_rrank) PK_CHECK(quiesce(c));
        dev_free(c->d_rrank);
        PK_CHECK(dev_alloc(&c->d_rrank, (size_t)c->cand_cap));
        c->rrank_cap = c->cand_cap;
    }
    if (c->eager_cap != M || !c->d_packed) {
        if (c->d_packed) PK_CHECK(quiesce(c));
        dev_free(c->d_perm); dev_free(c->d_packed);
        PK_CHECK(dev_alloc(&c->d_perm, (size_t)M));
        PK_CHECK(dev_alloc(&c->d_packed, (size_t)M * 28 + 16));
        c->eager_cap = M;
    }
    return PK_OK;
}

extern "C" int pk_chrom_features(pk_chrom* c, uint8_t* keep, float* fea32, double* fea64, int64_t capacity) {
    if (!c) { pk_set_error("pk_chrom_features: NULL handle"); return PK_EINVAL; }
    if (!c->has_candidates) { pk_set_error("pk_chrom_features: find_candidates not called"); return PK_ESTATE; }
    PK_CUDA(cudaSetDevice(c->device));
    if (!c->n_cand_known) PK_CHECK(settle_candidates(c));
    if (capacity < c->n_cand) { pk_set_error("pk_chrom_features: capacity too small"); return PK_ECAPACITY; }
    if (c->n_cand == 0) return PK_OK;
    PK_CHECK(reset_score_state(c));
    PK_CHECK(ensure_feature_buffer(c));
    double* d64 = nullptr;
    if (fea64) PK_CHECK(dev_alloc(&d64, (size_t)c->n_cand * c->F));
    int r = pk_launch_features(c, d64);
    if (r == PK_OK) {
        cudaError_t e = cudaSuccess;
        if (keep) e = cudaMemcpyAsync(keep, c->d_keep, (size_t)c->n_cand, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && fea32) e = cudaMemcpyAsync(fea32, c->d_fea32, (size_t)c->n_cand * c->F * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && fea64) e = cudaMemcpyAsync(fea64, d64, (size_t)c->n_cand * c->F * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { pk_set_error("pk_chrom_features: %s", cudaGetErrorString(e)); r = PK_ECUDA; }
    }
    dev_free(d64);
    c->has_scores = false;
    return r;
}

// Window features at caller-supplied pixels: the training-set extraction of trainUtils.buildmatrix
// (trainUtils.py:12-44). The pixels replace the handle's candidate list.
extern "C" int pk_chrom_features_at(pk_chrom* c, const int32_t* x, const int32_t* y, int64_t n, uint8_t* keep, float* fea32,
                                    double* fea64) {
    if (!c || n < 0 || (n > 0 && (!x || !y))) { pk_set_error("pk_chrom_features_at: bad argument"); return PK_EINVAL; }
    if (!c->has_pixels || !c->has_expected) { pk_set_error("pk_chrom_features_at: pixels and expected curve needed"); return PK_ESTATE; }
    std::vector<int32_t> hx((size_t)n), hd((size_t)n), hr((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t d = y[i] - x[i];
        if (x[i] < 0 || y[i] >= c->n || d < 0) { pk_set_error("pk_chrom_features_at: pixel %lld (%d, %d) is not in the upper triangle", (long long)i, x[i], y[i]); return PK_EINVAL; }
        // every cell of the window must lie on a stored diagonal with an expected value
        if (d + 2 * c->w >= c->ND - 1) { pk_set_error("pk_chrom_features_at: pixel %lld at distance %d needs upper >= %d", (long long)i, d, d + 1); return PK_EINVAL; }
        hx[(size_t)i] = x[i]; hd[(size_t)i] = d;
    }
    PK_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    c->has_candidates = false; c->has_scores = false;
    if (n == 0) return PK_OK;
    PK_CHECK(reserve_candidates(c, n));
    PK_CUDA(cudaMemcpyAsync(c->d_cx, hx.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(c->d_cd, hd.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(c->d_crank, hr.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    long long nc[2] = {(long long)n, (long long)n};
    PK_CUDA(cudaMemcpyAsync(c->d_ncand, nc, sizeof nc, cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaStreamSynchronize(s));           // hx, hd, hr, nc are locals
    c->n_cand = n; c->n_cand_all = n; c->n_cand_known = true;
    PK_CHECK(reset_score_state(c));
    PK_CHECK(ensure_feature_buffer(c));
    double* d64 = nullptr;
    if (fea64) PK_CHECK(dev_alloc(&d64, (size_t)n * c->F));
    int r = pk_launch_features(c, d64);
    if (r == PK_OK) {
        cudaError_t e = cudaSuccess;
        if (keep) e = cudaMemcpyAsync(keep, c->d_keep, (size_t)n, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && fea32) e = cudaMemcpyAsync(fea32, c->d_fea32, (size_t)n * c->F * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && fea64) e = cudaMemcpyAsync(fea64, d64, (size_t)n * c->F * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { pk_set_error("pk_chrom_features_at: %s", cudaGetErrorString(e)); r = PK_ECUDA; }
    }
    dev_free(d64);
    return r;
}

// Parity tap of the product kernel: the float32 feature rows k_score_fused builds in shared memory
// (scoreUtils.py:70-93 -> what predict_proba reads), spilled to global memory by the kernel itself.
// Rows of rejected candidates are zero. The separate feature kernel (pk_chrom_features) is a
// different code path (plain IEEE divisions); this one is what pk_chrom_score runs.
extern "C" int pk_chrom_fused_features(pk_chrom* c, pk_forest* f, uint8_t* keep, float* fea32, int64_t capacity) {
    if (!c || !f) { pk_set_error("pk_chrom_fused_features: NULL handle"); return PK_EINVAL; }
    if (!c->has_candidates) { pk_set_error("pk_chrom_fused_features: find_candidates not called"); return PK_ESTATE; }
    if (f->device != c->device || f->n_features != c->F) { pk_set_error("pk_chrom_fused_features: forest does not match the handle"); return PK_EINVAL; }
    if (!pk_fused_supported(c->w, f->n_trees)) { pk_set_error("pk_chrom_fused_features: no fused kernel for width %d", c->w); return PK_EUNSUPPORTED; }
    PK_CUDA(cudaSetDevice(c->device));
    if (!c->n_cand_known) PK_CHECK(settle_candidates(c));
    if (capacity < c->n_cand) { pk_set_error("pk_chrom_fused_features: capacity too small"); return PK_ECAPACITY; }
    c->has_scores = false;
    if (c->n_cand == 0) return PK_OK;
    PK_CHECK(reset_score_state(c));
    PK_CHECK(ensure_feature_buffer(c));
    PK_CUDA(cudaMemsetAsync(c->d_fea32, 0, (size_t)c->n_cand * c->F * sizeof(float), c->stream));
    const int variant = g_tune_fused > 1 ? g_tune_fused - 1 : 0;
    PK_CHECK(pk_launch_fused(c, f, variant, -1.0, 0, g_tune_cf, g_tune_tma, c->d_fea32));
    if (keep) PK_CUDA(cudaMemcpyAsync(keep, c->d_keep, (size_t)c->n_cand, cudaMemcpyDeviceToHost, c->stream));
    if (fea32) PK_CUDA(cudaMemcpyAsync(fea32, c->d_fea32, (size_t)c->n_cand * c->F * 4, cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaStreamSynchronize(c->stream));
    return PK_OK;
}

static int run_score(pk_chrom* c, pk_forest* f, double min_prob) {
    cudaStream_t s = c->stream;
    PK_CHECK(reset_score_state(c));
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[7], s));
    bool fused = g_tune_fused != 0 && pk_fused_supported(c->w, f->n_trees);
    if (fused) {
        // features stay in shared memory; stage [4] reports the fused kernel, [5] is 0
        const int variant = g_tune_fused > 1 ? g_tune_fused - 1 : 0;
        const double thre = g_tune_prune ? min_prob : -1.0;
        int r;
        if (c->use_score_stream) {
            // The fused kernel may run on a second stream (pk_chrom_set_score_stream): the handle's
            // own stream waits for it, so everything queued on the handle later is ordered behind it.
            if (!c->ev_x) PK_CUDA(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
            PK_CUDA(cudaEventRecord(c->ev_x, s));
            PK_CUDA(cudaStreamWaitEvent(c->score_stream, c->ev_x, 0));
            c->stream = c->score_stream;
            r = pk_launch_fused(c, f, variant, thre, g_tune_reserve, g_tune_cf, g_tune_tma, nullptr);
            c->stream = s;
            PK_CUDA(cudaEventRecord(c->ev_x, c->score_stream));
            PK_CUDA(cudaStreamWaitEvent(s, c->ev_x, 0));
        } else {
            r = pk_launch_fused(c, f, variant, thre, 0, g_tune_cf, g_tune_tma, nullptr);
        }
        if (r == PK_EUNSUPPORTED) fused = false;     // shapes the fused kernel has no room for: the two-kernel path below
        else PK_CHECK(r);
    }
    if (fused) {
        if (c->timing) PK_CUDA(cudaEventRecord(c->ev[8], s));
    } else {
        if (!c->n_cand_known) {
            // the separate kernels size their launch from the host's candidate count; if the scan had
            // to be redone with larger buffers, the scoring state is sized again as well
            const int64_t cap0 = c->cand_cap;
            PK_CHECK(settle_candidates(c));
            if (c->cand_cap != cap0) PK_CHECK(reset_score_state(c));
        }
        PK_CHECK(ensure_feature_buffer(c));
        PK_CHECK(pk_launch_features(c, nullptr));
        if (c->timing) PK_CUDA(cudaEventRecord(c->ev[8], s));
        PK_CHECK(pk_launch_forest(f, c->d_fea32, c->d_keep, c->n_cand, nullptr, c->d_prob, s));
    }
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[9], s));
    PK_CHECK(pk_launch_emit(c, min_prob));
    // order + pack the records now, behind the scoring pass: later the SMs belong to the next chromosome
    c->counts_valid = false;
    PK_CHECK(pk_launch_sort_records_eager(c, c->eager_cap));
    c->eager_valid = true;
    if (c->timing) PK_CUDA(cudaEventRecord(c->ev[10], s));
    return PK_OK;
}

extern "C" int pk_chrom_set_score_stream(pk_chrom* c, void* stream) {
    if (!c) { pk_set_error("pk_chrom_set_score_stream: NULL handle"); return PK_EINVAL; }
    c->score_stream = (cudaStream_t)stream;
    c->use_score_stream = stream != nullptr;
    return PK_OK;
}

extern "C" int pk_chrom_score(pk_chrom* c, pk_forest* f, double min_prob) {
    if (!c || !f) { pk_set_error("pk_chrom_score: NULL handle"); return PK_EINVAL; }
    if (!c->has_candidates) { pk_set_error("pk_chrom_score: find_candidates not called"); return PK_ESTATE; }
    if (f->device != c->device) { pk_set_error("pk_chrom_score: forest lives on device %d, chromosome on %d", f->device, c->device); return PK_EINVAL; }
    if (f->n_features != c->F) { pk_set_error("pk_chrom_score: forest has %d features, window has %d", f->n_features, c->F); return PK_EINVAL; }
    PK_CUDA(cudaSetDevice(c->device));
    c->last_forest = f; c->last_thre = min_prob;
    PK_CHECK(run_score(c, f, min_prob));
    c->has_scores = true;
    return PK_OK;
}

extern "C" int pk_chrom_result_count(pk_chrom* c, int64_t* n_records, int64_t* n_candidates, int64_t* n_windows) {
    if (!c) { pk_set_error("pk_chrom_result_count: NULL handle"); return PK_EINVAL; }
    if (!c->has_scores) { pk_set_error("pk_chrom_result_count: score not called"); return PK_ESTATE; }
    PK_CUDA(cudaSetDevice(c->device));
    if (!c->n_cand_known) {
        // first host look at this scan: if a device-side capacity was exceeded the scan was
        // redone by settle_candidates and the scoring has to be replayed on the full list
        int32_t flags[4];
        PK_CHECK(read_flags(c, flags));
        const bool overflow = flags[0] != 0 || (flags[3] & 2) != 0;
        PK_CHECK(settle_candidates(c));
        if (overflow) PK_CHECK(run_score(c, c->last_forest, c->last_thre));
    }
    if (!c->counts_valid) {
        PK_CUDA(cudaMemcpyAsync(c->h_counts, c->d_counters, sizeof c->h_counts, cudaMemcpyDeviceToHost, c->stream));
        PK_CUDA(cudaStreamSynchronize(c->stream));
        c->counts_valid = true;
    }
    if (n_records) *n_records = (int64_t)c->h_counts[0];
    if (n_candidates) *n_candidates = c->n_cand;
    if (n_windows) *n_windows = (int64_t)c->h_counts[1];
    return PK_OK;
}

extern "C" int pk_chrom_batch_windows(pk_chrom* c, int64_t* out, int64_t capacity, int64_t* n_batches) {
    if (!c || !n_batches) { pk_set_error("pk_chrom_batch_windows: bad argument"); return PK_EINVAL; }
    if (!c->has_scores) { pk_set_error("pk_chrom_batch_windows: score not called"); return PK_ESTATE; }
    PK_CHECK(pk_chrom_result_count(c, nullptr, nullptr, nullptr));
    c->n_batches = (c->n_cand_all + PK_BATCH - 1) / PK_BATCH;
    *n_batches = c->n_batches;
    if (!out) return PK_OK;
    if (capacity < c->n_batches) { pk_set_error("pk_chrom_batch_windows: capacity too small"); return PK_ECAPACITY; }
    std::vector<int32_t> h((size_t)std::max<int64_t>(c->n_batches, 1));
    PK_CUDA(cudaMemcpyAsync(h.data(), c->d_batch_win, (size_t)c->n_batches * 4, cudaMemcpyDeviceToHost, c->stream));
    PK_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < c->n_batches; ++i) out[i] = h[(size_t)i];
    return PK_OK;
}

extern "C" int pk_chrom_fetch_results(pk_chrom* c, int32_t* out_x, int32_t* out_y, double* out_prob, double* out_val,
                                      int32_t* out_batch, int64_t capacity, int mem) {
    if (!c) { pk_set_error("pk_chrom_fetch_results: NULL handle"); return PK_EINVAL; }
    int64_t n = 0;
    PK_CHECK(pk_chrom_result_count(c, &n, nullptr, nullptr));
    if (capacity < n) { pk_set_error("pk_chrom_fetch_results: capacity %lld < %lld records", (long long)capacity, (long long)n); return PK_ECAPACITY; }
    if (n == 0) return PK_OK;
    cudaStream_t s = c->stream;
    if (mem == PK_MEM_DEVICE) {
        if (out_x) PK_CUDA(cudaMemcpyAsync(out_x, c->d_rx, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        if (out_y) PK_CUDA(cudaMemcpyAsync(out_y, c->d_ry, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        if (out_prob) PK_CUDA(cudaMemcpyAsync(out_prob, c->d_rp, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        if (out_val) PK_CUDA(cudaMemcpyAsync(out_val, c->d_rv, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        if (out_batch) PK_CUDA(cudaMemcpyAsync(out_batch, c->d_rb, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        return PK_OK;
    }
    // sort by (x, y) on the device, pack, one copy into pinned staging, split on the host
    const long long off_f64 = ((12 * n + 7) / 8) * 8;
    const size_t packed_bytes = (size_t)off_f64 + 16 * (size_t)n;
    if (c->eager_valid && n <= c->eager_cap) {
        // already sorted and packed behind the scoring pass: only the copy is left
        if (packed_bytes > c->h_stage_bytes) {
            hstage_release(c->h_stage, c->h_stage_bytes);
            c->h_stage = nullptr; c->h_stage_bytes = 0;
            PK_CHECK(hstage_acquire(&c->h_stage, &c->h_stage_bytes, packed_bytes));
        }
        PK_CUDA(cudaMemcpyAsync(c->h_stage, c->d_packed, packed_bytes, cudaMemcpyDeviceToHost, s));
        PK_CUDA(cudaStreamSynchronize(s));
        const int32_t* hi = reinterpret_cast<const int32_t*>(c->h_stage);
        const double* hd = reinterpret_cast<const double*>(c->h_stage + off_f64);
        if (out_x) memcpy(out_x, hi, (size_t)n * 4);
        if (out_y) memcpy(out_y, hi + n, (size_t)n * 4);
        if (out_batch) memcpy(out_batch, hi + 2 * n, (size_t)n * 4);
        if (out_prob) memcpy(out_prob, hd, (size_t)n * 8);
        if (out_val) memcpy(out_val, hd + n, (size_t)n * 8);
        return PK_OK;
    }
    // every record is a candidate and the ordering buffers cover cand_cap records: not reachable
    pk_set_error("pk_chrom_fetch_results: %lld records exceed the ordering buffers (%lld)", (long long)n, (long long)c->eager_cap);
    return PK_ECAPACITY;
}

extern "C" int pk_chrom_stage_ms(pk_chrom* c, float* out_ms) {
    if (!c || !out_ms) { pk_set_error("pk_chrom_stage_ms: bad argument"); return PK_EINVAL; }
    PK_CUDA(cudaSetDevice(c->device));
    PK_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 8; ++i) out_ms[i] = 0.f;
    out_ms[0] = ev_ms(c, 0, 1);
    out_ms[1] = ev_ms(c, 1, 2);
    out_ms[2] = ev_ms(c, 3, 4);
    out_ms[3] = ev_ms(c, 5, 6);
    if (c->has_scores) {
        out_ms[4] = ev_ms(c, 7, 8);
        out_ms[5] = ev_ms(c, 8, 9);
        out_ms[6] = ev_ms(c, 9, 10);
    }
    return PK_OK;
}

// ---------------------------------------------------------------------------
// engine: score_genome's loop (score_genome.py:46-84) as a persistent per-device pipeline.
// Streams, chromosome handles and pinned staging are created once and reused from pass to pass;
// pk_engine_submit queues a unit (a chromosome, or a band row tile of one) without ever waiting
// for the device; pk_engine_collect waits for the units in order and moves every unit's records
// with one copy each into one pinned block. Uploads and the short stages of the units run on a
// ring of high-priority streams, the scoring kernels back to back on one ordinary stream.
// ---------------------------------------------------------------------------
struct pk_engine_slot {
    pk_chrom* c = nullptr;
    int64_t tag = 0;
    int32_t row_begin = 0, row_end = 0;
    cudaEvent_t done = nullptr;         // recorded behind the copy of the head block
    size_t head_off = 0;                // in h_head
    int64_t n_rec = 0, n_cand = 0, n_cand_all = 0, n_win = 0;
    int32_t n_bins = 0;
    size_t res_off = 0;                 // in h_res
};

struct pk_engine {
    int device = 0;
    pk_forest* forest = nullptr;
    int32_t w = 0, lower = 0, upper = 0;
    std::vector<cudaStream_t> streams;
    cudaStream_t score_stream = nullptr;
    std::vector<pk_engine_slot> units;  // queued since the last collect
    std::vector<pk_chrom*> idle;        // handles kept for reuse
    std::vector<cudaEvent_t> events;    // one per slot index, created on demand
    unsigned char* h_head = nullptr; size_t h_head_bytes = 0, h_head_used = 0;
    unsigned char* h_res = nullptr; size_t h_res_bytes = 0, h_res_used = 0;
    int64_t submitted = 0;
    int in_flight = 0, max_in_flight = 0;   // units holding a chromosome handle; more than that and the oldest is retired first
    bool collected = true;              // results of the last collect are still exposed
};

static const size_t PK_ENGINE_HEAD_BYTES = 4 << 20;
static const size_t PK_ENGINE_IDLE_MAX = 64;

extern "C" int pk_engine_create(int device, pk_forest* f, int32_t width, int32_t lower, int32_t upper, int depth,
                                pk_engine** out) {
    if (!out || !f || depth < 1 || depth > 32) { pk_set_error("pk_engine_create: bad argument"); return PK_EINVAL; }
    if (f->device != device) { pk_set_error("pk_engine_create: forest lives on device %d", f->device); return PK_EINVAL; }
    if ((2 * width + 1) * (2 * width + 1) != f->n_features) {
        pk_set_error("pk_engine_create: forest has %d features, width %d needs %d", f->n_features, width, (2 * width + 1) * (2 * width + 1));
        return PK_EINVAL;
    }
    PK_CUDA(cudaSetDevice(device));
    pk_engine* e = new pk_engine();
    e->device = device; e->forest = f; e->w = width; e->lower = lower; e->upper = upper;
    e->max_in_flight = 2 * depth;
    int lo = 0, hi = 0;                       // numerically lower = higher priority
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    for (int i = 0; i < depth; ++i) {
        cudaStream_t s = nullptr;
        if (cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi) != cudaSuccess) { pk_engine_destroy(e); pk_set_error("pk_engine_create: stream"); return PK_ECUDA; }
        e->streams.push_back(s);
    }
    if (cudaStreamCreateWithPriority(&e->score_stream, cudaStreamNonBlocking, lo) != cudaSuccess) { pk_engine_destroy(e); pk_set_error("pk_engine_create: stream"); return PK_ECUDA; }
    if (cudaMallocHost((void**)&e->h_head, PK_ENGINE_HEAD_BYTES) != cudaSuccess) { pk_engine_destroy(e); pk_set_error("pk_engine_create: pinned memory"); return PK_ENOMEM; }
    e->h_head_bytes = PK_ENGINE_HEAD_BYTES;
    *out = e;
    return PK_OK;
}

static void engine_drain(pk_engine* e) {
    for (auto s : e->streams) cudaStreamSynchronize(s);
    if (e->score_stream) cudaStreamSynchronize(e->score_stream);
}

// forget the queued units (after an error): wait for the device, keep the handles
extern "C" int pk_engine_reset(pk_engine* e) {
    if (!e) return PK_OK;
    cudaSetDevice(e->device);
    engine_drain(e);
    for (auto& u : e->units)
        if (u.c) pk_chrom_destroy(u.c);        // their state is unknown: do not reuse
    e->units.clear();
    e->h_head_used = 0; e->h_res_used = 0; e->in_flight = 0;
    e->collected = true;
    cudaGetLastError();
    return PK_OK;
}

extern "C" int pk_engine_destroy(pk_engine* e) {
    if (!e) return PK_OK;
    cudaSetDevice(e->device);
    pk_engine_reset(e);
    for (auto c : e->idle) pk_chrom_destroy(c);
    for (auto ev : e->events) if (ev) cudaEventDestroy(ev);
    for (auto s : e->streams) if (s) cudaStreamDestroy(s);
    if (e->score_stream) cudaStreamDestroy(e->score_stream);
    if (e->h_head) cudaFreeHost(e->h_head);
    if (e->h_res) cudaFreeHost(e->h_res);
    delete e;
    return PK_OK;
}

static int engine_retire(pk_engine* e, pk_engine_slot& u);

extern "C" int pk_engine_submit(pk_engine* e, const pk_unit* u) {
    if (!e || !u) { pk_set_error("pk_engine_submit: NULL argument"); return PK_EINVAL; }
    if (u->n_bins <= 0 || u->row_begin < 0 || u->row_end > u->n_bins || u->row_begin > u->row_end) {
        pk_set_error("pk_engine_submit: bad unit (n_bins %d, rows [%d, %d))", u->n_bins, u->row_begin, u->row_end);
        return PK_EINVAL;
    }
    PK_CUDA(cudaSetDevice(e->device));
    if (e->collected) {                      // first unit of a new pass: the previous results are released
        e->collected = false;
        e->h_head_used = 0;
        e->h_res_used = 0;
    }
    // bound the device memory of a long queue: beyond max_in_flight the oldest unit is finished first
    for (auto& q : e->units) {
        if (e->in_flight < e->max_in_flight) break;
        if (q.c) PK_CHECK(engine_retire(e, q));
    }
    const int balanced = u->weights != nullptr;
    pk_chrom* c = nullptr;
    for (size_t i = 0; i < e->idle.size(); ++i)
        if (e->idle[i]->n == u->n_bins && e->idle[i]->balanced == balanced) {
            c = e->idle[i];
            e->idle.erase(e->idle.begin() + (long)i);
            break;
        }
    cudaStream_t s = e->streams[(size_t)(e->submitted++ % (int64_t)e->streams.size())];
    if (!c) {
        PK_CHECK(pk_chrom_create(e->device, u->n_bins, e->w, e->lower, e->upper, balanced, (void*)s, &c));
        c->timing = false;
        c->score_stream = e->score_stream;
        c->use_score_stream = true;
    }
    c->stream = s;
    if (c->reuse_pending) {                  // the previous unit's record copy may still be on its old stream
        PK_CUDA(cudaStreamWaitEvent(s, c->ev_x, 0));
        c->reuse_pending = false;
    }
    const size_t hb = (c->head_bytes + 63) & ~(size_t)63;
    int r = PK_OK;
    if (e->h_head_used + hb > e->h_head_bytes) { pk_set_error("pk_engine_submit: too many units queued; collect first"); r = PK_ECAPACITY; }
    if (r == PK_OK) {
        switch (u->encoding) {
        case PK_ENC_COO: r = pk_chrom_upload_pixels(c, (const int32_t*)u->a, (const int32_t*)u->b, (const int32_t*)u->c, u->size, u->weights, PK_MEM_HOST | PK_PIXELS_SORTED); break;
        case PK_ENC_CSR32: r = pk_chrom_upload_csr(c, (const int64_t*)u->a, (const int32_t*)u->b, (const int32_t*)u->c, u->size, u->weights, PK_MEM_HOST); break;
        case PK_ENC_CSR16: r = pk_chrom_upload_csr16(c, (const int64_t*)u->a, (const uint16_t*)u->b, (const uint16_t*)u->c, u->size, u->weights, PK_MEM_HOST); break;
        case PK_ENC_ROWS: r = pk_chrom_upload_rows(c, u->a, u->size, u->weights, PK_MEM_HOST); break;
        default: pk_set_error("pk_engine_submit: unknown encoding %d", u->encoding); r = PK_EINVAL;
        }
    }
    if (r == PK_OK && u->poisson_weights) r = pk_chrom_set_poisson_weights(c, u->poisson_weights, PK_MEM_HOST);
    if (r == PK_OK) r = pk_chrom_fit_expected(c);
    if (r == PK_OK) r = pk_chrom_find_candidates(c, u->row_begin, u->row_end, nullptr);
    if (r == PK_OK) r = pk_chrom_score(c, e->forest, u->min_prob);
    pk_engine_slot slot;
    if (r == PK_OK) {
        const size_t idx = e->units.size();
        while (e->events.size() <= idx) {
            cudaEvent_t ev = nullptr;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { pk_set_error("pk_engine_submit: event"); r = PK_ECUDA; break; }
            e->events.push_back(ev);
        }
        if (r == PK_OK) {
            slot.c = c; slot.tag = u->tag; slot.row_begin = u->row_begin; slot.row_end = u->row_end;
            slot.done = e->events[idx];
            slot.head_off = e->h_head_used;
            cudaError_t ce = cudaMemcpyAsync(e->h_head + slot.head_off, c->d_head, c->head_bytes, cudaMemcpyDeviceToHost, s);
            if (ce == cudaSuccess) ce = cudaEventRecord(slot.done, s);
            if (ce != cudaSuccess) { pk_set_error("pk_engine_submit: %s", cudaGetErrorString(ce)); r = PK_ECUDA; }
        }
    }
    if (r != PK_OK) {
        cudaStreamSynchronize(s);
        if (c->score_stream) cudaStreamSynchronize(c->score_stream);
        pk_chrom_destroy(c);
        return r;
    }
    e->h_head_used += hb;
    e->units.push_back(slot);
    ++e->in_flight;
    return PK_OK;
}

static int engine_grow_results(pk_engine* e, size_t need) {
    if (need <= e->h_res_bytes) return PK_OK;
    engine_drain(e);                          // earlier record copies land in the old block first
    size_t want = std::max<size_t>(2 * need, 16 << 20);
    unsigned char* nb = nullptr;
    if (cudaMallocHost((void**)&nb, want) != cudaSuccess) { pk_set_error("pk_engine_collect: %zu bytes of pinned memory", want); cudaGetLastError(); return PK_ENOMEM; }
    if (e->h_res) { memcpy(nb, e->h_res, e->h_res_bytes); cudaFreeHost(e->h_res); }
    e->h_res = nb; e->h_res_bytes = want;
    return PK_OK;
}

// Finish the oldest unit still in flight: wait for its head block, queue the copy of its records into the
// pinned result block, hand its chromosome handle back for reuse (the next user's stream waits for that copy).
static int engine_retire(pk_engine* e, pk_engine_slot& u) {
    pk_chrom* c = u.c;
    PK_CUDA(cudaEventSynchronize(u.done));
    unsigned char* hd = e->h_head + u.head_off;
    const int32_t* flags = reinterpret_cast<const int32_t*>(hd);
    const long long* nc = reinterpret_cast<const long long*>(hd + 16);
    const unsigned long long* cnt = reinterpret_cast<const unsigned long long*>(hd + 32);
    const bool overflow = flags[0] != 0 || (flags[3] & 2) != 0 || nc[0] > c->cand_cap;
    const bool trouble = overflow || (flags[2] & 3) != 0 || (c->declared_sorted && (flags[3] & 1));
    if (trouble) {
        // a device-side capacity was exceeded (the scan and the scoring are replayed) or the pass failed:
        // the step-by-step entry points sort it out, then the head block is read again
        PK_CHECK(pk_chrom_result_count(c, nullptr, nullptr, nullptr));
        PK_CUDA(cudaMemcpyAsync(hd, c->d_head, c->head_bytes, cudaMemcpyDeviceToHost, c->stream));
        PK_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        c->n_cand = nc[0]; c->n_cand_all = nc[1]; c->n_cand_known = true;
        memcpy(c->h_counts, cnt, sizeof c->h_counts);
        c->counts_valid = true;
    }
    u.n_rec = (int64_t)c->h_counts[0];
    u.n_cand = c->n_cand; u.n_cand_all = c->n_cand_all; u.n_win = (int64_t)c->h_counts[1]; u.n_bins = c->n;
    if (u.n_rec > c->eager_cap) { pk_set_error("pk_engine: %lld records exceed the ordering buffers", (long long)u.n_rec); return PK_ECAPACITY; }
    const size_t off_f64 = (size_t)((12 * u.n_rec + 7) / 8) * 8;
    const size_t bytes = off_f64 + 16 * (size_t)u.n_rec;
    u.res_off = e->h_res_used;
    e->h_res_used = (e->h_res_used + bytes + 63) & ~(size_t)63;
    PK_CHECK(engine_grow_results(e, e->h_res_used));
    if (u.n_rec > 0)
        PK_CUDA(cudaMemcpyAsync(e->h_res + u.res_off, c->d_packed, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (!c->ev_x) PK_CUDA(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
    PK_CUDA(cudaEventRecord(c->ev_x, c->stream));        // the next unit on this handle starts behind the copy
    c->reuse_pending = true;
    if (e->idle.size() < PK_ENGINE_IDLE_MAX) e->idle.push_back(c);
    else { pk_chrom_destroy(c); }
    u.c = nullptr;
    --e->in_flight;
    return PK_OK;
}

extern "C" int pk_engine_collect(pk_engine* e, pk_unit_result* out, int64_t capacity, int64_t* n_units) {
    if (!e || !n_units) { pk_set_error("pk_engine_collect: NULL argument"); return PK_EINVAL; }
    *n_units = (int64_t)e->units.size();
    if (!out) return PK_OK;
    if (capacity < (int64_t)e->units.size()) { pk_set_error("pk_engine_collect: capacity %lld < %zu units", (long long)capacity, e->units.size()); return PK_ECAPACITY; }
    PK_CUDA(cudaSetDevice(e->device));
    for (auto& u : e->units)
        if (u.c) PK_CHECK(engine_retire(e, u));
    engine_drain(e);
    PK_CUDA(cudaGetLastError());
    int64_t i = 0;
    for (auto& u : e->units) {
        pk_unit_result& r = out[i++];
        const unsigned char* hd = e->h_head + u.head_off;
        const unsigned char* base = e->h_res + u.res_off;
        const size_t off_f64 = (size_t)((12 * u.n_rec + 7) / 8) * 8;
        r.tag = u.tag; r.n_bins = u.n_bins; r.row_begin = u.row_begin; r.row_end = u.row_end;
        r.whole = (u.row_begin == 0 && u.row_end == u.n_bins) ? 1 : 0;
        r.n_records = u.n_rec; r.n_candidates = u.n_cand; r.n_windows = u.n_win;
        r.n_batches = (u.n_cand_all + PK_BATCH - 1) / PK_BATCH;
        r.x = reinterpret_cast<const int32_t*>(base);
        r.y = r.x + u.n_rec;
        r.batch = r.x + 2 * u.n_rec;
        r.prob = reinterpret_cast<const double*>(base + off_f64);
        r.value = r.prob + u.n_rec;
        r.batch_windows = reinterpret_cast<const int32_t*>(hd + 64);
    }
    e->units.clear();
    e->collected = true;
    return PK_OK;
}
