// Microbenchmark: cycles per DFMA for ONE warp per SM with 1..8 independent chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void chains(double* out, long long* cyc, int iters, double m, double c) {
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = threadIdx.x * 1e-9 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (s == 1.2345) out[0] = s;
}

template <int CH>
void run(int warps_per_sm, int active_lanes) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 1024);
    const int iters = 2000;
    chains<CH><<<148, dim3(active_lanes < 32 && warps_per_sm == 1 ? active_lanes : 32 * warps_per_sm), 0>>>(out, cyc, iters, 0.999999, 1e-9);
    chains<CH><<<148, dim3(active_lanes < 32 && warps_per_sm == 1 ? active_lanes : 32 * warps_per_sm), 0>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double per = (double)h[0] / (iters * 8.0 * CH);
    printf("chains=%d warps/SM=%d lanes=%d : %.2f cycles per DFMA per warp (%.2f per chain step)\n", CH, warps_per_sm, active_lanes, per, per * CH);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1>(1, 32); run<2>(1, 32); run<4>(1, 32); run<8>(1, 32); run<16>(1, 32);
    run<8>(1, 16); run<8>(1, 8);
    run<8>(4, 32); run<8>(8, 32); run<8>(16, 32);
    run<2>(8, 32); run<1>(8, 32);
    return 0;
}
