// Micro-benchmark: how fast does one SM fetch small 2-D TMA boxes (a window's parallelogram of the band:
// (4w+1) diagonals x (2w+2) int32 cells) at scattered coordinates, and do out-of-bounds rows/columns
// (negative diagonals, rows past the tensor) read as zero?  Decides whether phase A1 of k_score_fused can
// be a cp.async.bulk.tensor.2d per window instead of 242 predicated 4-byte loads per window pair.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_box tma_box.cu && ./tma_box
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_box_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

constexpr int ROWS = 21, COLS = 12, SLOT = 1024;      // box of 21 x 12 int32 = 1008 bytes in a 1 KB slot

// NG groups of 128 threads; each group repeatedly fetches PB boxes (one lane each), waits, checks them.
template <int NG, int PB>
__global__ void __launch_bounds__(NG * 128, 1) k_boxes(const __grid_constant__ CUtensorMap map, int n, int ND, long long pitch, int iters,
                                                       unsigned long long* bad, unsigned long long* sink, int check, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[NG];
    const int tid = threadIdx.x, grp = tid / 128, gtid = tid % 128;
    unsigned char* buf = smem + (size_t)grp * PB * SLOT;
    if (gtid == 0) mbar_init(&bars[grp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned long long acc = 0, nbad = 0;
    uint32_t rng = 1234567u + 977u * (blockIdx.x * NG + grp);
    for (int it = 0; it < iters; ++it) {
        // coordinates of this take (same in every thread of the group)
        int bx[PB], bd[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i) {
            rng = rng * 1664525u + 1013904223u; bx[i] = (int)((rng >> 8) % (unsigned)(n - 400)) + 5;
            rng = rng * 1664525u + 1013904223u; bd[i] = (int)((rng >> 8) % 300u) + 4;      // some d < 10: negative rows
            if (mode & 1) bx[i] = (bx[i] & ~3) + 5;        // box starts on a 16-byte boundary
            if (mode & 2) bd[i] = bd[i] < 10 ? 10 : (bd[i] > 280 ? 280 : bd[i]);     // no out-of-bounds rows
        }
        if (gtid == 0) mbar_expect_tx(&bars[grp], PB * ROWS * COLS * 4);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
        for (int i = 0; i < PB; ++i)
            if (gtid == ((mode & 4) ? (i % 4) * 32 + i / 4 : (mode & 8) ? (i % 4) * 32 : i))      // who issues: one warp, lanes of four warps, lane 0 of four warps
                tma_box_2d(buf + i * SLOT, &map, bx[i] - 5, bd[i] - 10, &bars[grp]);
        mbar_wait(&bars[grp], it & 1);
        // consume: every thread reads a few cells; optionally verify all of them
        const int32_t* cells = reinterpret_cast<const int32_t*>(buf);
        if (!check) acc += (unsigned)cells[(gtid % PB) * (SLOT / 4) + gtid];
        else for (int q = gtid; q < PB * ROWS * COLS; q += 128) {
            const int i = q / (ROWS * COLS), r = (q / COLS) % ROWS, c = q % COLS;
            const int32_t v = cells[i * (SLOT / 4) + r * COLS + c];
            acc += (unsigned)v;
            if (check) {
                const int dd = bd[i] - 10 + r, pp = bx[i] - 5 + c;
                const int32_t want = (dd >= 0 && dd < ND - 1 && pp >= 0 && pp < (int)pitch) ? (dd * 100003 + pp) : 0;
                if (v != want) ++nbad;
            }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    }
    if (nbad) atomicAdd(bad, nbad);
    if (acc == 0xdeadbeefull) *sink = acc;
}

template <int NG, int PB>
static void run(const CUtensorMap& map, int n, int ND, long long pitch, unsigned long long* d_bad, unsigned long long* d_sink, int mode) {
    const size_t smem = (size_t)NG * PB * SLOT;
    CK(cudaFuncSetAttribute(k_boxes<NG, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaMemset(d_bad, 0, 8));
    k_boxes<NG, PB><<<148, NG * 128, smem>>>(map, n, ND, pitch, 20, d_bad, d_sink, 1, mode);
    CK(cudaDeviceSynchronize());
    unsigned long long bad = 0;
    CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
    const int iters = 400;
    CK(cudaEventRecord(a));
    k_boxes<NG, PB><<<148, NG * 128, smem>>>(map, n, ND, pitch, iters, d_bad, d_sink, 0, mode);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double boxes_per_sm = (double)iters * NG * PB;
    printf("mode %d NG=%d PB=%2d: mismatching cells %llu; %.1f us for %.0f boxes per SM -> %.1f ns (%.0f cycles at 1.965 GHz) per box per SM, %.2f GB/s of smem fill per SM\n",
           mode, NG, PB, bad, ms * 1e3, boxes_per_sm, ms * 1e6 / boxes_per_sm, ms * 1e6 / boxes_per_sm * 1.965,
           boxes_per_sm * ROWS * COLS * 4 / (ms * 1e-3) / 1e9);
}


// ---- second experiment: row-major band B2[r][o] (o = c - r, row pitch P2 = 4k + 1) viewed through a SKEWED tensor
// map T[j][i] = base + j * (P2 - 1) + i, which is the dense matrix cell (row j, column i): a window is then the plain
// box rows x-w..x+w, columns y-w..y+w (start aligned down to 4 columns: 16 columns x 11 rows), half the rows of the
// diagonal-major box. Plus the two weight vectors as 1-D boxes of 12 doubles from an even index.
constexpr int R2 = 11, C2 = 16, SLOT2 = 1024;
template <int NG, int PB>
__global__ void __launch_bounds__(NG * 128, 1) k_boxes2(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap wmap,
                                                        int n, int ND, int P2, int iters, unsigned long long* bad, unsigned long long* sink, int check) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[NG];
    const int tid = threadIdx.x, grp = tid / 128, gtid = tid % 128;
    unsigned char* buf = smem + (size_t)grp * PB * SLOT2;
    if (gtid == 0) mbar_init(&bars[grp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned long long acc = 0, nbad = 0;
    uint32_t rng = 1234567u + 977u * (blockIdx.x * NG + grp);
    for (int it = 0; it < iters; ++it) {
        int bx[PB], bd[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i) {
            rng = rng * 1664525u + 1013904223u; bx[i] = (int)((rng >> 8) % (unsigned)(n - 400)) + 5;
            rng = rng * 1664525u + 1013904223u; bd[i] = (int)((rng >> 8) % 290u) + 10;
        }
        if (gtid == 0) mbar_expect_tx(&bars[grp], PB * (R2 * C2 * 4 + 2 * 96));
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
        for (int i = 0; i < PB; ++i)
            if (gtid == i) {
                const int y0 = bx[i] + bd[i] - 5;
                tma_box_2d(buf + i * SLOT2, &map, y0 & ~3, bx[i] - 5, &bars[grp]);
                asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                             ::"r"(smem_u32(buf + i * SLOT2 + 768)), "l"(&wmap), "r"(smem_u32(&bars[grp])), "r"((bx[i] - 5) & ~1) : "memory");
                asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                             ::"r"(smem_u32(buf + i * SLOT2 + 896)), "l"(&wmap), "r"(smem_u32(&bars[grp])), "r"(y0 & ~1) : "memory");
            }
        mbar_wait(&bars[grp], it & 1);
        if (!check) acc += (unsigned)reinterpret_cast<const int32_t*>(buf + (gtid % PB) * SLOT2)[gtid];
        else for (int q = gtid; q < PB * 121; q += 128) {
            const int i = q / 121, a = (q / 11) % 11, b = q % 11;
            const int y0 = bx[i] + bd[i] - 5, off = y0 & 3;
            const int32_t v = reinterpret_cast<const int32_t*>(buf + i * SLOT2)[a * C2 + off + b];
            const double wr = reinterpret_cast<const double*>(buf + i * SLOT2 + 768)[((bx[i] - 5) & 1) + a];
            const double wc = reinterpret_cast<const double*>(buf + i * SLOT2 + 896)[(y0 & 1) + b];
            acc += (unsigned)v + (unsigned long long)wr + (unsigned long long)wc;
            if (check) {
                const int r = bx[i] - 5 + a, c = y0 + b, o = c - r;
                const int32_t want = (o >= 0 && o < ND - 1) ? (r * 331 + o) : 0;
                if (v != want || wr != (double)r * 0.5 || wc != (double)c * 0.5) ++nbad;
            }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    }
    if (nbad) atomicAdd(bad, nbad);
    if (acc == 0xdeadbeefull) *sink = acc;
}

template <int NG, int PB>
static void run2(const CUtensorMap& map, const CUtensorMap& wmap, int n, int ND, int P2, unsigned long long* d_bad, unsigned long long* d_sink) {
    const size_t smem = (size_t)NG * PB * SLOT2;
    CK(cudaFuncSetAttribute(k_boxes2<NG, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaMemset(d_bad, 0, 8));
    k_boxes2<NG, PB><<<148, NG * 128, smem>>>(map, wmap, n, ND, P2, 20, d_bad, d_sink, 1);
    CK(cudaDeviceSynchronize());
    unsigned long long bad = 0;
    CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
    const int iters = 400;
    CK(cudaEventRecord(a));
    k_boxes2<NG, PB><<<148, NG * 128, smem>>>(map, wmap, n, ND, P2, iters, d_bad, d_sink, 0);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double boxes_per_sm = (double)iters * NG * PB;
    printf("skewed row-major NG=%d PB=%2d: mismatching cells %llu; %.1f us for %.0f windows per SM -> %.1f ns (%.0f cycles at 1.965 GHz) per window per SM\n",
           NG, PB, bad, ms * 1e3, boxes_per_sm, ms * 1e6 / boxes_per_sm, ms * 1e6 / boxes_per_sm * 1.965);
}

static int skewed(EncodeTiled enc, unsigned long long* d_bad, unsigned long long* d_sink) {
    const int n = 24900, ND = 311, P2 = 313;      // P2 = 4k + 1 >= ND
    std::vector<int32_t> h((size_t)n * P2 + 64, 0);
    for (int r = 0; r < n; ++r)
        for (int o = 0; o < ND - 1; ++o) h[(size_t)r * P2 + o] = r * 331 + o;
    std::vector<double> hw(n + 16);
    for (int i = 0; i < n + 16; ++i) hw[i] = 0.5 * i;
    int32_t* d_b2; double* d_w;
    CK(cudaMalloc(&d_b2, h.size() * 4)); CK(cudaMemcpy(d_b2, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_w, hw.size() * 8)); CK(cudaMemcpy(d_w, hw.data(), hw.size() * 8, cudaMemcpyHostToDevice));
    CUtensorMap map, wmap;
    const cuuint64_t dims[2] = {(cuuint64_t)n + 16, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)(P2 - 1) * 4};
    const cuuint32_t box[2] = {C2, R2}, estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d_b2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("skewed map (dim0 %d > stride/4 %d): cuTensorMapEncodeTiled -> %d\n", n + 16, P2 - 1, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const cuuint64_t wd[1] = {(cuuint64_t)n + 16};
    const cuuint32_t wb[1] = {12}, we[1] = {1};
    const cuuint64_t ws[1] = {0};
    r = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1, d_w, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("weight map: cuTensorMapEncodeTiled -> %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    run2<4, 10>(map, wmap, n, ND, P2, d_bad, d_sink);
    run2<4, 20>(map, wmap, n, ND, P2, d_bad, d_sink);
    run2<2, 23>(map, wmap, n, ND, P2, d_bad, d_sink);
    return 0;
}

int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int n = 24900, ND = 311;
    const long long pitch = (n + 31) / 32 * 32;
    std::vector<int32_t> h((size_t)ND * pitch);
    for (int d = 0; d < ND; ++d)
        for (long long p = 0; p < pitch; ++p) h[(size_t)d * pitch + p] = d * 100003 + (int)p;
    int32_t* d_band;
    CK(cudaMalloc(&d_band, h.size() * 4));
    CK(cudaMemcpy(d_band, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    unsigned long long *d_bad, *d_sink;
    CK(cudaMalloc(&d_bad, 8)); CK(cudaMalloc(&d_sink, 8));

    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
    if (!enc || qres != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)(ND - 1)};      // the last stored diagonal is trimmed
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    const cuuint32_t box[2] = {COLS, ROWS}, estr[2] = {1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d_band, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled -> %d\n", (int)r); return 1; }
    if (mode == 9) { skewed(enc, d_bad, d_sink); mode = 3; }

    run<4, 11>(map, n, ND, pitch, d_bad, d_sink, mode);
    run<4, 16>(map, n, ND, pitch, d_bad, d_sink, mode);
    run<4, 20>(map, n, ND, pitch, d_bad, d_sink, mode);
    run<2, 23>(map, n, ND, pitch, d_bad, d_sink, mode);
    run<1, 32>(map, n, ND, pitch, d_bad, d_sink, mode);
    run<4, 4>(map, n, ND, pitch, d_bad, d_sink, mode);
    return 0;
}
