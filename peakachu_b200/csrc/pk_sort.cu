// peakachu_b200: order the emitted records by (x, y) -- the order prob_csr.nonzero()
// yields (scoreUtils.py:130) -- and pack them for one device-to-host copy.
// Not on the hot path: a few thousand records per chromosome. The radix sort is CUB's.
#include <cub/device/device_radix_sort.cuh>

#include "pk_common.cuh"

__global__ void __launch_bounds__(256) k_record_keys(const int32_t* __restrict__ rx, const int32_t* __restrict__ ry,
                                                     long long n, unsigned long long* __restrict__ keys,
                                                     uint32_t* __restrict__ idx) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    keys[i] = ((unsigned long long)(uint32_t)rx[i] << 32) | (uint32_t)ry[i];
    idx[i] = (uint32_t)i;
}

// packed layout: x[n] | y[n] | batch[n] (int32), then prob[n] | value[n] (float64, 8-byte aligned)
__global__ void __launch_bounds__(256) k_record_gather(const uint32_t* __restrict__ order, long long n,
                                                       const int32_t* __restrict__ rx, const int32_t* __restrict__ ry,
                                                       const int32_t* __restrict__ rb, const double* __restrict__ rp,
                                                       const double* __restrict__ rv, unsigned char* __restrict__ packed,
                                                       long long off_f64) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = order[i];
    int32_t* pi = reinterpret_cast<int32_t*>(packed);
    double* pd = reinterpret_cast<double*>(packed + off_f64);
    pi[i] = rx[j]; pi[n + i] = ry[j]; pi[2 * n + i] = rb[j];
    pd[i] = rp[j]; pd[n + i] = rv[j];
}

size_t pk_sort_temp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return bytes;
}

// keys_in/out [n] u64, idx_in/out [n] u32, temp: pk_sort_temp_bytes(n); packed: 12n rounded to 8, + 16n bytes
int pk_launch_sort_records(pk_chrom* c, long long n, unsigned long long* keys_in, unsigned long long* keys_out,
                           uint32_t* idx_in, uint32_t* idx_out, void* temp, size_t temp_bytes, unsigned char* packed,
                           long long off_f64, int key_bits) {
    if (n == 0) return PK_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    k_record_keys<<<grid, 256, 0, c->stream>>>(c->d_rx, c->d_ry, n, keys_in, idx_in);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, idx_out, (int)n, 0, key_bits,
                                            c->stream));
    k_record_gather<<<grid, 256, 0, c->stream>>>(idx_out, n, c->d_rx, c->d_ry, c->d_rb, c->d_rp, c->d_rv, packed, off_f64);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

// ---------------------------------------------------------------------------
// Eager variant, queued right behind the scoring pass: the record count is still on the
// device, so the sort covers a fixed capacity M (unused slots carry the largest key) and
// the pack reads the count itself. The host later copies packed_bytes(count) -- no kernel
// has to be scheduled at fetch time, when other chromosomes' kernels own the SMs.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_record_keys32(const int32_t* __restrict__ rx, const int32_t* __restrict__ ry,
                                                       const unsigned long long* __restrict__ counters, long long M,
                                                       uint32_t nd, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= M) return;
    const long long n = (long long)counters[0];
    keys[i] = i < n ? (uint32_t)rx[i] * nd + (uint32_t)(ry[i] - rx[i]) : 0xFFFFFFFFu;
    idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_record_gather_dev(const uint32_t* __restrict__ order, const unsigned long long* __restrict__ counters,
                                                           long long M, const int32_t* __restrict__ rx, const int32_t* __restrict__ ry,
                                                           const int32_t* __restrict__ rb, const double* __restrict__ rp,
                                                           const double* __restrict__ rv, unsigned char* __restrict__ packed) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long n = (long long)counters[0];
    if (i >= n || n > M) return;
    const long long off_f64 = ((12 * n + 7) / 8) * 8;
    const uint32_t j = order[i];
    int32_t* pi = reinterpret_cast<int32_t*>(packed);
    double* pd = reinterpret_cast<double*>(packed + off_f64);
    pi[i] = rx[j]; pi[n + i] = ry[j]; pi[2 * n + i] = rb[j];
    pd[i] = rp[j]; pd[n + i] = rv[j];
}

size_t pk_sort32_temp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n);
    return bytes;
}

int pk_launch_sort_records_eager(pk_chrom* c, long long M, int key_bits) {
    const unsigned grid = (unsigned)((M + 255) / 256);
    k_record_keys32<<<grid, 256, 0, c->stream>>>(c->d_rx, c->d_ry, c->d_counters, M, (uint32_t)c->ND, c->d_sk0, c->d_si0);
    PK_CUDA(cudaGetLastError());
    size_t tb = c->stemp_bytes;
    PK_CUDA(cub::DeviceRadixSort::SortPairs(c->d_stemp, tb, c->d_sk0, c->d_sk1, c->d_si0, c->d_si1, (int)M, 0, key_bits, c->stream));
    k_record_gather_dev<<<grid, 256, 0, c->stream>>>(c->d_si1, c->d_counters, M, c->d_rx, c->d_ry, c->d_rb, c->d_rp, c->d_rv,
                                                     c->d_packed);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}
