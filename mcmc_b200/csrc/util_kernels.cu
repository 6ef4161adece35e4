// Small device entry points used by tests and diagnostics: evaluate a registered target
// functor at given points, and dump the raw Philox variate stream.
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"

namespace mcmcb200
{

template <class T, int EPL, bool STRICT> __global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) eval_kernel(const EvalLaunch a)
{
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pt = (long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    if (pt >= a.n_points) return;
    const int d = a.d;
    const int dpad = (d + 1) & ~1;
    const WarpCtx w{lane, d, smem + (size_t)warp * dpad};
    double x[EPL], g[EPL];
    load_vec<EPL>(a.x + pt * d, d, lane, x);
    double v;
    if (a.grad) {
        v = T::template eval<EPL, STRICT, true, true>(a.tdata, w, x, g);
        store_vec<EPL>(a.grad + pt * d, d, lane, g);
    } else {
        v = T::template eval<EPL, STRICT, true, false>(a.tdata, w, x, g);
    }
    if (lane == 0) a.value[pt] = v;
}

template <class T, int EPL> static int eval_one(const EvalLaunch& a)
{
    const long long blocks = (a.n_points + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int dpad = (a.d + 1) & ~1;
    const size_t smem = (size_t)WARPS_PER_BLOCK * dpad * sizeof(double);
    if (a.strict) eval_kernel<T, EPL, true><<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    else eval_kernel<T, EPL, false><<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T> static int eval_target(const EvalLaunch& a)
{
    switch (epl_for_dim(a.d)) {
    MCMCB200_EPL_CASE(2, (eval_one<T, 2>(a)))
    MCMCB200_EPL_CASE(4, (eval_one<T, 4>(a)))
    MCMCB200_EPL_CASE(8, (eval_one<T, 8>(a)))
    MCMCB200_EPL_CASE(16, (eval_one<T, 16>(a)))
    default: set_error("target_eval: n_dim=%d unsupported", a.d); return MCMCB200_ERR_UNSUPPORTED;
    }
}

int MCMCB200_SLICED(launch_target_eval)(const EvalLaunch& a)
{
    switch (a.target_id) {
#define X(ID, TYPE) \
    case ID: return eval_target<TYPE>(a);
        MCMCB200_FOREACH_TARGET(X)
#undef X
    default:
#ifndef MCMCB200_USER_TARGET_TYPE
        if (user_target_has(USER_LAUNCH_EVAL, a.target_id)) {
            EvalLaunch b = a;
            b.target_id = MCMCB200_TARGET_USER;
            return user_target_launch(USER_LAUNCH_EVAL, a.target_id, &b);
        }
#endif
        set_error("target_eval: unknown target id %d", a.target_id);
        return MCMCB200_ERR_UNKNOWN_TARGET;
    }
}

#ifndef MCMCB200_USER_TARGET_TYPE   // the raw-stream dump and the peak probe belong to the library only
// fp64 peak probe: 8 independent DFMA chains per thread, 8 warps per SM sub-partition — the denominator of the
// compute-bound rooflines (C3 / C4 / C5), measured in the same process as the kernels it is compared with.
__global__ void __launch_bounds__(1024) fp64_peak_kernel(double* out, int iters, double a, double b)
{
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = threadIdx.x * 1e-9 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fma(v[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 123.456) out[0] = s;   // never true: keeps the loop alive
}

int launch_fp64_peak(double* scratch_dev, cudaStream_t stream, double* tflops_out)
{
    int dev = 0, sms = 0;
    MCMCB200_CUDA_TRY(cudaGetDevice(&dev));
    MCMCB200_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaEvent_t e0, e1;
    MCMCB200_CUDA_TRY(cudaEventCreate(&e0));
    MCMCB200_CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 20000;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        MCMCB200_CUDA_TRY(cudaEventRecord(e0, stream));
        fp64_peak_kernel<<<sms * 2, 1024, 0, stream>>>(scratch_dev, iters, 0.999999, 1e-7);
        MCMCB200_CUDA_TRY(cudaEventRecord(e1, stream));
        MCMCB200_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        MCMCB200_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_out = 2.0 * 8.0 * iters * 1024.0 * sms * 2 / (best * 1e-3) / 1e12;
    return MCMCB200_OK;
}

template <int EPL> __global__ void philox_stream_kernel(const RngArgs r, long long chain, long long draw, int d, int n_unif, double* out)
{
    __shared__ double2 log_tab[RNG_TAB_DOUBLE2];
    build_rng_tables(log_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    ChainRng<RNG_PHILOX> rng;
    rng.init(r, 0, chain);
    double z[EPL];
    rng.template normals<EPL, false>(r, draw, d, lane, log_tab, z);
    store_vec<EPL>(out, d, lane, z);
    for (int k = 0; k < n_unif; ++k) {
        const double u = rng.uniform(r, draw, k);
        if (lane == 0) out[d + k] = u;
    }
}

int launch_philox_stream(unsigned long long seed, long long chain, long long draw, int d, int n_unif, double* out_dev,
                         cudaStream_t stream)
{
    RngArgs r;
    r.mode = RNG_PHILOX;
    rng_set_key(r, seed);
    r.tape = nullptr;
    r.tape_stride = 0;
    r.err_flag = nullptr;
    switch (epl_for_dim(d)) {
    case 2: philox_stream_kernel<2><<<1, 32, 0, stream>>>(r, chain, draw, d, n_unif, out_dev); break;
    case 4: philox_stream_kernel<4><<<1, 32, 0, stream>>>(r, chain, draw, d, n_unif, out_dev); break;
    case 8: philox_stream_kernel<8><<<1, 32, 0, stream>>>(r, chain, draw, d, n_unif, out_dev); break;
    case 16: philox_stream_kernel<16><<<1, 32, 0, stream>>>(r, chain, draw, d, n_unif, out_dev); break;
    default: set_error("philox_stream: n_dim=%d unsupported", d); return MCMCB200_ERR_UNSUPPORTED;
    }
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}
#endif

}  // namespace mcmcb200
