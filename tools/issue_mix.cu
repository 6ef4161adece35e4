// Micro-benchmark: do integer/logic instructions issue "for free" next to fp64 instructions on an SMSP,
// or does a DFMA occupy the dispatch port for 2 cycles?  Times D DFMAs + I integer ops per iteration.
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NI> __global__ void mix(double* out, unsigned* iout, int iters, double a, double b, unsigned m)
{
    double v[8];
    unsigned w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x * 1e-9 + i; w[i] = threadIdx.x * 2654435761u + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < ND) v[i] = fma(v[i], a, b);
                if (i < NI) w[i] = (w[i] * m) ^ (w[i] >> 7);   // IMAD + SHF + LOP3 -> ~2-3 int ops
            }
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += v[i]; t ^= w[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int ND, int NI> void run(const char* name, double* out, unsigned* iout, int sms)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, threads = 256, blocks = sms * 4;  // 32 warps / SM = 8 per SMSP
    mix<ND, NI><<<blocks, threads>>>(out, iout, 10, 1.0000001, 1e-9, 2654435761u);
    cudaEventRecord(e0);
    mix<ND, NI><<<blocks, threads>>>(out, iout, iters, 1.0000001, 1e-9, 2654435761u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    // cycles per iteration-row (ND dfma + NI int groups) per warp per SMSP, assuming ~1.9 GHz actual
    double warp_rows = (double)blocks * threads / 32 * iters * 4 / (sms * 4);
    printf("%-22s ND=%d NI=%d: %.3f ms  -> %.2f ns per (row) per SMSP-warp-slot\n", name, ND, NI, ms, ms * 1e6 / warp_rows);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; unsigned* iout;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 4 * 256);
    cudaMalloc(&iout, sizeof(unsigned) * p.multiProcessorCount * 4 * 256);
    run<8, 0>("dfma only", out, iout, p.multiProcessorCount);
    run<0, 8>("int only", out, iout, p.multiProcessorCount);
    run<8, 8>("dfma + int", out, iout, p.multiProcessorCount);
    run<4, 8>("half dfma + int", out, iout, p.multiProcessorCount);
    run<8, 4>("dfma + half int", out, iout, p.multiProcessorCount);
    return 0;
}
