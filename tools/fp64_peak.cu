// Micro-benchmark: sustained fp64 FMA rate and dependent-issue latency on the device.
// Gives the fp64 co-roof that DESIGN.md / bench.py quote next to the HBM roofline.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP> __global__ void dfma_throughput(double* out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_latency(double* out, long long* cyc, int iters, double a, double b)
{
    double v = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) v = fma(v, a, b);
    const long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * 16 * 256);
    cudaMalloc(&cyc, sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps_per_sm : {4, 8, 16, 32, 64}) {
        const int threads = 256;
        const int blocks = sms * warps_per_sm * 32 / threads;
        dfma_throughput<8><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        dfma_throughput<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fmas = (double)blocks * threads * 8.0 * iters;
        printf("warps/SM=%2d ILP=8: %.2f TFLOP/s fp64 (%.1f DFMA/clk/SM at %d MHz nominal)\n", warps_per_sm, 2 * fmas / ms * 1e-9,
               fmas / (ms * 1e-3) / sms / (p.clockRate * 1e3), p.clockRate / 1000);
    }
    dfma_latency<<<1, 32>>>(out, cyc, 10000, 1.0000001, 1e-9);
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)h / 10000);
    printf("SMs=%d clock=%d MHz name=%s\n", sms, p.clockRate / 1000, p.name);
    return 0;
}
