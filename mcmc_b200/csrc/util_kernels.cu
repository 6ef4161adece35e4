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
    case 2: return eval_one<T, 2>(a);
    case 4: return eval_one<T, 4>(a);
    case 8: return eval_one<T, 8>(a);
    case 16: return eval_one<T, 16>(a);
    default: set_error("target_eval: n_dim=%d unsupported", a.d); return MCMCB200_ERR_UNSUPPORTED;
    }
}

int launch_target_eval(const EvalLaunch& a)
{
    switch (a.target_id) {
#define X(ID, TYPE) \
    case ID: return eval_target<TYPE>(a);
        MCMCB200_FOREACH_TARGET(X)
#undef X
    default: set_error("target_eval: unknown target id %d", a.target_id); return MCMCB200_ERR_UNKNOWN_TARGET;
    }
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

}  // namespace mcmcb200
