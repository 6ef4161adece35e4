// Developer tool: time the C2 headline kernel (HMC d=128 iso-Gaussian, 4096 chains x 1100 draws, L=10, eps=0.1,
// Philox, FAST arithmetic) as a standalone executable, so that several compile-time variants (-DMCMCB200_V_*) can be
// built side by side and compared in ONE gpurun call.  Not part of the product; the kernel source is the product's own
// (#include of csrc/hmc.cu with the launch dispatch compiled out).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -DMCMCB200_KERNEL_ONLY \
//        [-DMCMCB200_V_...] tools/c2_variants.cu -o tools/bin/c2_<name>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../mcmc_b200/csrc/hmc.cu"

namespace mcmcb200
{
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}
int epl_for_dim(int d) { return d <= 64 ? 2 : d <= 128 ? 4 : d <= 256 ? 8 : d <= 512 ? 16 : 0; }
}  // namespace mcmcb200

using namespace mcmcb200;

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e = (x);                                                                \
        if (e != cudaSuccess) {                                                             \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                        \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

#ifndef C2_LS
#define C2_LS 10
#endif
#ifndef C2_UNR
#define C2_UNR 1
#endif
#ifndef C2_EPL
#define C2_EPL 4
#endif

int main(int argc, char** argv)
{
    const int C = argc > 1 ? atoi(argv[1]) : 4096, d = 32 * C2_EPL, nb = 100, nk = argc > 3 ? atoi(argv[3]) : 1000, reps = argc > 2 ? atoi(argv[2]) : 6;
    std::vector<double> x0((size_t)C * d);
    for (int c = 0; c < C; ++c)
        for (int j = 0; j < d; ++j) x0[(size_t)c * d + j] = sin(0.37 * c + 0.11 * j);
    double *dx0, *ddraws;
    long long* dacc;
    CK(cudaMalloc(&dx0, x0.size() * 8));
    CK(cudaMalloc(&ddraws, (size_t)C * nk * d * 8));
    CK(cudaMalloc(&dacc, (size_t)C * 8));
    CK(cudaMemcpy(dx0, x0.data(), x0.size() * 8, cudaMemcpyHostToDevice));
    HmcLaunch a{};
    a.n_chains = C; a.d = d; a.target_id = 0; a.tdata = nullptr; a.x0 = dx0; a.broadcast_x0 = 0; a.chain_offset = 0;
    a.rng.mode = RNG_PHILOX; rng_set_key(a.rng, 12345ull); a.rng.tape = nullptr; a.rng.tape_stride = 0;
    a.draws = ddraws; a.logp = nullptr; a.n_accept = dacc; a.stream = 0; a.strict = false; a.lb = a.ub = nullptr;
    a.n_burnin = nb; a.n_keep = nk; a.n_leap = 10; a.eps = 0.1; a.S_cm = nullptr; a.Minv_cm = nullptr;
#ifdef C2_OLD
    auto kern = hmc_kernel<IsoGauss, C2_EPL, false, false, RNG_PHILOX, true, false>;
#else
    auto kern = hmc_pipe_kernel<IsoGauss, C2_EPL, C2_LS, C2_UNR>;
#endif
    const size_t smem = (size_t)WARPS_PER_BLOCK * d * sizeof(double);
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS_PER_BLOCK * 32, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        kern<<<(C + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) { best = ms < best ? ms : best; sum += ms; }
    }
    // checksum over a slice: draws of chains 0..63, plus all accept counts
    std::vector<double> h((size_t)(C < 64 ? C : 64) * nk * d);
    std::vector<long long> acc(C);
    CK(cudaMemcpy(h.data(), ddraws, h.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(acc.data(), dacc, (size_t)C * 8, cudaMemcpyDeviceToHost));
    double cs = 0, cs2 = 0;
    for (size_t i = 0; i < h.size(); ++i) { cs += h[i]; cs2 += h[i] * h[i]; }
    long long na = 0;
    for (int c = 0; c < C; ++c) na += acc[c];
    printf("%-28s regs=%d occ=%d  best %.4f ms  mean %.4f ms  -> %.3e draws/s  roofline %.3f | sum %.10e var %.10f acc %lld last %.17g\n", argv[0],
           fa.numRegs, occ, best, sum / (reps - 1), (double)C * (nb + nk) / best * 1e3, (double)C * (nb + nk) * 16.0 * d / (best * 1e-3) / 6454.6e9, cs,
           cs2 / h.size(), na, h[h.size() - 1]);
    return 0;
}
