// peakachu_b200: order the emitted records by (x, y) -- the order prob_csr.nonzero()
// yields (scoreUtils.py:130) -- and pack them for one device-to-host copy.
#include "pk_common.cuh"
#include "pk_device.cuh"

// ---------------------------------------------------------------------------
// Queued right behind the scoring pass, when the record count is still on the device, for any
// number of records (--minimum-prob 0 emits every kept window). k_emit has counted the records of every row x (rowcnt) and given each record its
// arrival number within the row (rrank). Three short kernels then order the records by (x, y):
//   k_row_offsets   exclusive scan of rowcnt (one CTA)
//   k_record_place  perm[rowoff[x] + rrank] = record; clears rowcnt for the next pass
//   k_record_pack   rank of a record inside its row = records of the row with a smaller y
//                   (rows hold a handful of records); writes the packed layout
// The host later copies packed_bytes(count): no kernel has to be scheduled at fetch time, when
// other chromosomes' kernels own the SMs.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_row_offsets(const int32_t* __restrict__ rowcnt, int n, int32_t* __restrict__ rowoff) {
    __shared__ uint32_t s_warp[33];
    pk_cta_scan_1024(rowcnt, (long long)n, rowoff, s_warp);
}

__global__ void __launch_bounds__(256) k_record_place(const unsigned long long* __restrict__ counters, long long M,
                                                      const int32_t* __restrict__ rx, const int32_t* __restrict__ rrank,
                                                      const int32_t* __restrict__ rowoff, int32_t* __restrict__ rowcnt,
                                                      uint32_t* __restrict__ perm) {
    const long long n = (long long)counters[0];
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const int x = rx[i], r = rrank[i];
        if (r == 0) rowcnt[x] = 0;                          // the scan has consumed it
        if (n <= M) perm[rowoff[x] + r] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(256) k_record_pack(const unsigned long long* __restrict__ counters, long long M,
                                                     const uint32_t* __restrict__ perm, const int32_t* __restrict__ rowoff,
                                                     const int32_t* __restrict__ rx, const int32_t* __restrict__ ry,
                                                     const int32_t* __restrict__ rb, const double* __restrict__ rp,
                                                     const double* __restrict__ rv, unsigned char* __restrict__ packed) {
    const long long n = (long long)counters[0];
    if (n > M) return;
    const long long off_f64 = ((12 * n + 7) / 8) * 8;
    int32_t* pi = reinterpret_cast<int32_t*>(packed);
    double* pd = reinterpret_cast<double*>(packed + off_f64);
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const int x = rx[i], y = ry[i];
        const int a = rowoff[x], b = rowoff[x + 1];
        int rank = 0;
        for (int j = a; j < b; ++j) rank += ry[perm[j]] < y;
        const long long o = a + rank;
        pi[o] = x; pi[n + o] = y; pi[2 * n + o] = rb[i];
        pd[o] = rp[i]; pd[n + o] = rv[i];
    }
}

int pk_launch_sort_records_eager(pk_chrom* c, long long M) {
    k_row_offsets<<<1, 1024, 0, c->stream>>>(c->d_rowcnt, c->n, c->d_rowoff);
    PK_CUDA(cudaGetLastError());
    k_record_place<<<148, 256, 0, c->stream>>>(c->d_counters, M, c->d_rx, c->d_rrank, c->d_rowoff, c->d_rowcnt, c->d_perm);
    PK_CUDA(cudaGetLastError());
    k_record_pack<<<148, 256, 0, c->stream>>>(c->d_counters, M, c->d_perm, c->d_rowoff, c->d_rx, c->d_ry, c->d_rb, c->d_rp,
                                              c->d_rv, c->d_packed);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}
