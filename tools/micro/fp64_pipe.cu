// Micro-benchmark: FP64 pipe issue interval and dependent latency on the device it runs on.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int OP>
__global__ void k_dp(double* out, double a, double b, int iters, long long* cycles) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a + c + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) x[c] = __fma_rn(x[c], a, b);
            else if (OP == 1) x[c] = __dadd_rn(x[c], b);
            else x[c] = __dmul_rn(x[c], a);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS, int OP>
void run(const char* name, int threads) {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k_dp<CHAINS, OP><<<148, threads>>>(out, 1.0000001, 1e-9, iters, cyc);
    k_dp<CHAINS, OP><<<148, threads>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_instr_per_smsp = (double)iters * CHAINS * (threads / 32) / 4.0;
    printf("%-5s chains=%d threads=%4d: %8lld cycles, %.2f cycles per warp-instr per SMSP, %.1f cycles per iteration\n", name, CHAINS,
           threads, h, h / warp_instr_per_smsp, (double)h / iters);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1, 0>("dfma", 32);      // dependent latency
    run<1, 1>("dadd", 32);
    run<1, 2>("dmul", 32);
    run<2, 0>("dfma", 32);
    run<4, 0>("dfma", 32);
    run<8, 0>("dfma", 32);
    run<8, 0>("dfma", 128);     // 1 warp per SMSP
    run<8, 0>("dfma", 256);
    run<8, 0>("dfma", 512);
    run<8, 0>("dfma", 1024);
    run<8, 1>("dadd", 512);
    run<8, 2>("dmul", 512);
    run<2, 0>("dfma", 512);
    run<1, 0>("dfma", 512);
    return 0;
}
