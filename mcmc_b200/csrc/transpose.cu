// [n_chains][n_keep][d] -> [n_chains][d][n_keep]: the column-major n_keep x n_dim Mat_t the reference returns per chain
// (src/hmc.cpp:138), produced where the draws already are.  32 x 32 tiles through padded shared memory, both
// sides coalesced (256-byte segments per warp row); HBM-bound: 16 bytes per element moved.
#include "engine.h"
#include "transpose.h"

namespace mcmcb200
{

__global__ void __launch_bounds__(256) transpose_draws_kernel(const double* __restrict__ in, double* __restrict__ out, long long n_keep, int d)
{
    __shared__ double tile[32][33];
    const long long chain = blockIdx.z;
    const double* src = in + chain * n_keep * d;
    double* dst = out + chain * n_keep * d;
    const long long t0 = (long long)blockIdx.y * 32;
    const int j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const long long t = t0 + r;
        const int j = j0 + tx;
        tile[r][tx] = (t < n_keep && j < d) ? src[t * d + j] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r;
        const long long t = t0 + tx;
        if (j < d && t < n_keep) dst[(long long)j * n_keep + t] = tile[tx][r];
    }
}

int launch_transpose_draws(const double* in, double* out, long long n_chains, long long n_keep, int d, cudaStream_t stream)
{
    if (n_chains <= 0 || n_keep <= 0) return MCMCB200_OK;
    const long long ty = (n_keep + 31) / 32;
    if (ty > 65535) { set_error("transpose: n_keep too large for one grid"); return MCMCB200_ERR_UNSUPPORTED; }
    for (long long c0 = 0; c0 < n_chains; c0 += 65535) {
        const long long nc = n_chains - c0 < 65535 ? n_chains - c0 : 65535;
        dim3 grid((unsigned)((d + 31) / 32), (unsigned)ty, (unsigned)nc);
        transpose_draws_kernel<<<grid, 256, 0, stream>>>(in + c0 * n_keep * d, out + c0 * n_keep * d, n_keep, d);
    }
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

}  // namespace mcmcb200
