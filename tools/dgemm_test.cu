// Stand-alone check + timing of the DMMA fp64 GEMM used by the chain-batched MALA path (mala_wide.cu).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../mcmc_b200/csrc/mala_wide.cu"
namespace mcmcb200 { void set_error(const char* fmt, ...) { (void)fmt; } int epl_for_dim(int) { return 0; } }
using namespace mcmcb200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void naive(const double* Y, const double* A, double* C, int M, int d)
{
    const int r = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    double s = 0;
    for (int k = 0; k < d; ++k) s = fma(Y[(size_t)r * d + k], A[(size_t)k * d + c], s);
    C[(size_t)r * d + c] = s;
}

int main(int argc, char** argv)
{
    const int M = argc > 1 ? atoi(argv[1]) : 16384, d = argc > 2 ? atoi(argv[2]) : 1024;
    std::vector<double> hY((size_t)M * d), hA((size_t)d * d);
    srand(1);
    for (auto& v : hY) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < d; ++i) for (int j = 0; j <= i; ++j) { double v = rand() / (double)RAND_MAX - 0.5; hA[(size_t)i * d + j] = v; hA[(size_t)j * d + i] = v; }
    double *Y, *A, *C, *R;
    CK(cudaMalloc(&Y, hY.size() * 8)); CK(cudaMalloc(&A, hA.size() * 8)); CK(cudaMalloc(&C, hY.size() * 8)); CK(cudaMalloc(&R, hY.size() * 8));
    CK(cudaMemcpy(Y, hY.data(), hY.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
    const size_t gsmem = (size_t)G_STAGES * G_STAGE_DOUBLES * sizeof(double);
    printf("smem %zu bytes\n", gsmem);
    CK(cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
    const dim3 grid((d + GN - 1) / GN, (M + GM - 1) / GM);
    dgemm_dmma_kernel<<<grid, G_THREADS, gsmem>>>(Y, A, C, M, d);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    naive<<<dim3((d + 127) / 128, M), 128>>>(Y, A, R, M, d);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    std::vector<double> hC(hY.size()), hR(hY.size());
    CK(cudaMemcpy(hC.data(), C, hC.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hR.data(), R, hR.size() * 8, cudaMemcpyDeviceToHost));
    double worst = 0; for (size_t i = 0; i < hC.size(); ++i) worst = fmax(worst, fabs(hC[i] - hR[i]));
    printf("M=%d d=%d max |dmma - naive| = %.3e\n", M, d, worst);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) dgemm_dmma_kernel<<<grid, G_THREADS, gsmem>>>(Y, A, C, M, d);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) dgemm_dmma_kernel<<<grid, G_THREADS, gsmem>>>(Y, A, C, M, d);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("dgemm_dmma: %.3f ms -> %.2f TFLOP/s fp64\n", ms, 2.0 * M * d * d / ms * 1e-9);
    return 0;
}
