// fp64 tensor-core GEMM shared by the chain-batched sampler paths (defined in mala_wide.cu): C[M x d] = Y[M x d] * A[d x d],
// row-major, hand-written DMMA (mma.sync.m8n8k4.f64 — tcgen05 has no f64 kind), cp.async 3-stage pipeline.
#pragma once
#include <cuda_runtime.h>

namespace mcmcb200
{
int launch_dgemm_dmma(const double* Y, const double* Amat, double* Cout, long long M, int d, cudaStream_t stream);
}
