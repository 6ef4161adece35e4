// draws_out layout conversion on the device (transpose.cu): [c][t][j] chain-major rows -> [c][j][t], i.e. every chain's
// n_keep x n_dim matrix in column-major order — the reference's Mat_t.
#pragma once
#include <cuda_runtime.h>

namespace mcmcb200
{
int launch_transpose_draws(const double* in, double* out, long long n_chains, long long n_keep, int d, cudaStream_t stream);
}
