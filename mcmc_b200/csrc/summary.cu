// On-device summaries of draws_out (include/mcmc_b200_summary.h).
//
// chain_stats_kernel: one warp per (chain, 64-element block of the row): streams the chain's n_keep rows once
//   (each lane a 16-byte vector per row, 512 contiguous bytes per warp instruction, 8 rows in flight) and accumulates
//   sum (x - s) and sum (x - s)^2 with s = the chain's first kept draw, so large means do not cancel.  HBM-bound: the
//   draws are read exactly once (8 B per element), everything else is O(n_chains * n_dim).
// combine_kernel: per tile of 32 elements, 32 x 32 threads over (element, chain slice): pooled mean, pooled variance, R-hat.
#include <vector>

#include "engine.h"
#include "../../include/mcmc_b200_summary.h"

namespace mcmcb200
{

constexpr int SUM_WARPS = 4;
constexpr int SUM_UNROLL = 8;

__global__ void __launch_bounds__(SUM_WARPS * 32) chain_stats_kernel(const double* __restrict__ draws, long long n_chains, long long n_keep,
                                                                      int d, int blocks_per_row, double* __restrict__ cmean,
                                                                      double* __restrict__ cvar)
{
    const long long wid = (long long)blockIdx.x * SUM_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= n_chains * blocks_per_row) return;
    const long long chain = wid / blocks_per_row;
    const int j = (int)(wid % blocks_per_row) * 64 + 2 * lane;
    if (j >= d) return;
    const bool pair = (j + 1 < d);
    const bool vec = pair && ((d & 1) == 0) && ((reinterpret_cast<uintptr_t>(draws) & 15) == 0);
    const double* p = draws + (size_t)chain * (size_t)n_keep * d + j;
    auto ld = [&](long long t, double& a, double& b) {
        const double* q = p + (size_t)t * d;
        if (vec) {
            const double2 v = __ldcs(reinterpret_cast<const double2*>(q));   // streamed once: evict-first
            a = v.x; b = v.y;
        } else {
            a = __ldcs(q);
            b = pair ? __ldcs(q + 1) : 0.0;
        }
    };
    double s0 = 0.0, s1 = 0.0;
    if (n_keep > 0) ld(0, s0, s1);
    double a0 = 0.0, a1 = 0.0, q0 = 0.0, q1 = 0.0;
    long long t = 0;
    for (; t + SUM_UNROLL <= n_keep; t += SUM_UNROLL) {
        double v0[SUM_UNROLL], v1[SUM_UNROLL];
#pragma unroll
        for (int u = 0; u < SUM_UNROLL; ++u) ld(t + u, v0[u], v1[u]);
#pragma unroll
        for (int u = 0; u < SUM_UNROLL; ++u) {
            const double e0 = v0[u] - s0, e1 = v1[u] - s1;
            a0 += e0; a1 += e1;
            q0 = fma(e0, e0, q0); q1 = fma(e1, e1, q1);
        }
    }
    for (; t < n_keep; ++t) {
        double v0, v1;
        ld(t, v0, v1);
        const double e0 = v0 - s0, e1 = v1 - s1;
        a0 += e0; a1 += e1;
        q0 = fma(e0, e0, q0); q1 = fma(e1, e1, q1);
    }
    const double n = (double)n_keep;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const double m0 = a0 / n, m1 = a1 / n;   // mean of the shifted values
    double* om = cmean + (size_t)chain * d + j;
    double* ov = cvar + (size_t)chain * d + j;
    om[0] = n_keep > 0 ? s0 + m0 : nan;
    ov[0] = n_keep > 1 ? fmax(0.0, (q0 - a0 * m0) / (n - 1.0)) : nan;
    if (pair) {
        om[1] = n_keep > 0 ? s1 + m1 : nan;
        ov[1] = n_keep > 1 ? fmax(0.0, (q1 - a1 * m1) / (n - 1.0)) : nan;
    }
}

// one block per tile of 32 elements, 32 x 32 threads: thread (tx, ty) reduces chains ty, ty + 32, ... of element j = tile*32 + tx
// (coalesced across tx), then the 32 partials per element are combined in shared memory.  Sums of the chain means are shifted
// by chain 0's mean so that the between-chain variance does not cancel.
__global__ void __launch_bounds__(1024) combine_kernel(const double* __restrict__ cmean, const double* __restrict__ cvar, long long n_chains,
                                                       long long n_keep, int d, double* __restrict__ mean, double* __restrict__ var,
                                                       double* __restrict__ rhat)
{
    __shared__ double s1[32][33], s2[32][33], sw[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j = blockIdx.x * 32 + tx;
    double a1 = 0.0, a2 = 0.0, aw = 0.0;
    const double shift = (j < d) ? cmean[j] : 0.0;
    if (j < d)
        for (long long c = ty; c < n_chains; c += 32) {
            const double e = cmean[(size_t)c * d + j] - shift;
            a1 += e;
            a2 = fma(e, e, a2);
            aw += cvar[(size_t)c * d + j];
        }
    s1[ty][tx] = a1; s2[ty][tx] = a2; sw[ty][tx] = aw;
    __syncthreads();
    if (ty == 0 && j < d) {
        double t1 = 0.0, t2 = 0.0, tw = 0.0;
        for (int r = 0; r < 32; ++r) { t1 += s1[r][tx]; t2 += s2[r][tx]; tw += sw[r][tx]; }
        const double C = (double)n_chains, n = (double)n_keep;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        const double m = shift + t1 / C;
        const double sb = fmax(0.0, t2 - t1 * (t1 / C));          // sum_c (m_c - m)^2
        mean[j] = m;
        const double W = tw / C;                                   // mean within-chain variance
        const double Bn = n_chains > 1 ? sb / (C - 1.0) : nan;     // B/n: variance of the chain means
        // pooled sample variance of all C*n draws: [(n-1) sum_c s_c^2 + n sum_c (m_c - m)^2] / (C n - 1)
        var[j] = (n_keep > 1 || n_chains > 1) ? ((n_keep > 1 ? (n - 1.0) * tw : 0.0) + n * sb) / (C * n - 1.0) : nan;
        rhat[j] = (n_chains > 1 && n_keep > 1) ? sqrt(((n - 1.0) / n * W + Bn) / W) : nan;
    }
}

}  // namespace mcmcb200

using namespace mcmcb200;

extern "C" int mcmcb200_summarize_draws(const double* draws, int32_t draws_mem, int64_t n_chains, int64_t n_keep, int32_t n_dim,
                                        int32_t device, void* stream, mcmcb200_summary_t* out)
{
    if (!draws || !out || !out->mean || n_chains <= 0 || n_keep <= 0 || n_dim <= 0) {
        set_error("summarize_draws: draws, out->mean and positive sizes are required");
        return MCMCB200_ERR_INVALID_ARG;
    }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev <= 0) {
        set_error("no usable CUDA device (%s); mcmc_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return MCMCB200_ERR_CUDA;
    }
    int prev = 0;
    MCMCB200_CUDA_TRY(cudaGetDevice(&prev));
    const int dev = device < 0 ? prev : device;
    if (dev >= n_dev) { set_error("device %d out of range (%d visible)", dev, n_dev); return MCMCB200_ERR_INVALID_ARG; }
    struct Restore { int prev, dev; ~Restore() { if (prev != dev) cudaSetDevice(prev); } } restore{prev, dev};
    if (dev != prev) MCMCB200_CUDA_TRY(cudaSetDevice(dev));
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    // grow-only scratch of the calling host thread, per device (same policy as the engine's workspace)
    struct Scratch {
        void* p[16][3] = {};
        size_t n[16][3] = {};
        ~Scratch() { for (auto& d : p) for (void* q : d) if (q) cudaFree(q); }
        int get(int dev_, int slot, size_t bytes, void** out_)
        {
            if (n[dev_][slot] < bytes) {
                if (p[dev_][slot]) cudaFree(p[dev_][slot]);
                p[dev_][slot] = nullptr; n[dev_][slot] = 0;
                MCMCB200_CUDA_TRY(cudaMalloc(&p[dev_][slot], bytes));
                n[dev_][slot] = bytes;
            }
            *out_ = p[dev_][slot];
            return MCMCB200_OK;
        }
    };
    static thread_local Scratch scratch;
    if (dev >= 16) { set_error("device %d beyond the supported 16", dev); return MCMCB200_ERR_INVALID_ARG; }
    struct { void* p = nullptr; } up, stats, res;
    int rc;
    const size_t n_elem = (size_t)n_chains * (size_t)n_keep * (size_t)n_dim, n_cd = (size_t)n_chains * (size_t)n_dim;
    const double* d_draws = draws;
    if (draws_mem == MCMCB200_MEM_HOST) {
        if ((rc = scratch.get(dev, 0, n_elem * sizeof(double), &up.p))) return rc;
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(up.p, draws, n_elem * sizeof(double), cudaMemcpyHostToDevice, st));
        d_draws = static_cast<const double*>(up.p);
    }
    if ((rc = scratch.get(dev, 1, 2 * n_cd * sizeof(double), &stats.p))) return rc;
    if ((rc = scratch.get(dev, 2, 3 * (size_t)n_dim * sizeof(double), &res.p))) return rc;
    double* cmean = static_cast<double*>(stats.p);
    double* cvar = cmean + n_cd;
    double* r = static_cast<double*>(res.p);
    cudaEvent_t e0, e1;
    MCMCB200_CUDA_TRY(cudaEventCreate(&e0));
    MCMCB200_CUDA_TRY(cudaEventCreate(&e1));
    struct Ev { cudaEvent_t a, b; ~Ev() { cudaEventDestroy(a); cudaEventDestroy(b); } } evs{e0, e1};

    const int bpr = (n_dim + 63) / 64;
    const long long warps = (long long)n_chains * bpr;
    const long long blocks = (warps + SUM_WARPS - 1) / SUM_WARPS;
    if (blocks > 0x7fffffffll) { set_error("summarize_draws: problem too large for one launch"); return MCMCB200_ERR_UNSUPPORTED; }
    MCMCB200_CUDA_TRY(cudaEventRecord(e0, st));
    chain_stats_kernel<<<(unsigned)blocks, SUM_WARPS * 32, 0, st>>>(d_draws, n_chains, n_keep, n_dim, bpr, cmean, cvar);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    combine_kernel<<<(n_dim + 31) / 32, dim3(32, 32), 0, st>>>(cmean, cvar, n_chains, n_keep, n_dim, r, r + n_dim, r + 2 * n_dim);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    MCMCB200_CUDA_TRY(cudaEventRecord(e1, st));
    MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->mean, r, (size_t)n_dim * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out->var) MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->var, r + n_dim, (size_t)n_dim * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out->rhat) MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->rhat, r + 2 * n_dim, (size_t)n_dim * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out->chain_mean) MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->chain_mean, cmean, n_cd * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out->chain_var) MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->chain_var, cvar, n_cd * sizeof(double), cudaMemcpyDeviceToHost, st));
    MCMCB200_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    MCMCB200_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    out->kernel_ms = ms;
    return MCMCB200_OK;
}
