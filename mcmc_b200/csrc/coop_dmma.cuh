// Cooperative dense product of the NUTS kernel, producer side and fp64 tensor-core body (declared in warp.cuh, defined here so
// that tuning them rebuilds nuts.cu only).
#pragma once

#include "warp.cuh"

namespace mcmcb200
{

// Warp 0 (all 32 lanes call): a panel of ncols <= 32 matrix columns into a buffer whose columns are pstride doubles apart — one
// bulk copy by lane 0 when the buffer is dense (pstride == d); when it is padded (the DMMA path: pstride = d + 4 makes the
// 4 x 8 operand fragment reads bank-conflict-free) one bulk copy per column, issued by the 32 lanes in parallel (issued by a
// single thread the 256 copies of a round sat on the critical path: C4 got slower, not faster).
__device__ __forceinline__ void bulk_load_panel_cols(double* dst, const double* src, int ncols, int d, int pstride, unsigned long long* bar, int lane)
{
    if (pstride == d) {
        if (lane == 0) bulk_load_panel(dst, src, (unsigned)((size_t)ncols * d * sizeof(double)), bar);
        return;
    }
    if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" : : "r"(smem_u32(bar)), "r"((unsigned)((size_t)ncols * d * sizeof(double))) : "memory");
    __syncwarp();
    if (lane < ncols)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     : : "r"(smem_u32(dst + (size_t)lane * pstride)), "l"(src + (size_t)lane * d), "r"((unsigned)(d * sizeof(double))), "r"(smem_u32(bar)) : "memory");
}

// producer prologue (warp 0): the first two panels (their buffers are free: the previous product ended with a CTA barrier)
__device__ __forceinline__ void coop_panels_prologue(const double* __restrict__ A, int d, double* panels, int pstride, unsigned long long* mbar)
{
    if (threadIdx.x < 32) {
        const int np = (d + COOP_PANEL_COLS - 1) / COOP_PANEL_COLS;
        for (int k = 0; k < 2 && k < np; ++k) {
            const int j0 = k * COOP_PANEL_COLS;
            const int nc = (d - j0 < COOP_PANEL_COLS) ? d - j0 : COOP_PANEL_COLS;
            bulk_load_panel_cols(panels + (size_t)k * COOP_PANEL_COLS * pstride, A + (size_t)j0 * d, nc, d, pstride, mbar + k, (int)threadIdx.x);
        }
    }
}

// Tensor-core body (FAST arithmetic; d a multiple of 4, d <= 32*NW, NW == 8): Y[c][i] = sum_j X[c][j] A[j][i] for the 8 chains of
// the CTA as m8n8k4 fp64 MMAs — M = the 8 chains, N = 8 rows of the product, K = 4 matrix columns (tcgen05 has no f64 kind:
// DMMA is the fp64 tensor path of sm_100a).  Warp w owns rows 32w .. 32w+31 (four N tiles); per K step it loads ONE chain
// fragment (X[c = lane/4][j0 + lane%4], chains 2*dp + 4 doubles apart) and four matrix fragments (A[j0 + lane%4][i0 + 8n + lane/4]
// from the padded panel) and issues four DMMAs: 5 shared loads per 4 MMAs (= 32 multiply-adds per lane) where the scalar body
// needs 10 per 16 — the round stops being bound by instruction issue and shared-memory latency at 2 warps per sub-partition.
// The accumulation order is the tensor core's, so this body is used in FAST arithmetic only; STRICT keeps the scalar body.
__device__ __forceinline__ void coop_dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NW>
__device__ __forceinline__ void coop_gemv_body_dmma(const double* __restrict__ A, int d, int warp, int lane, double* __restrict__ base, int stride,
                                                    double* __restrict__ panels, int pstride, unsigned long long* mbar, unsigned& phase)
{
    const int half = stride >> 1;
    const int np = (d + COOP_PANEL_COLS - 1) / COOP_PANEL_COLS;
    const size_t pan_elems = (size_t)COOP_PANEL_COLS * pstride;
    const int g = lane >> 2, t = lane & 3;
    const int i0 = 32 * warp;
    const bool active = i0 < d;
    double acc[4][2];
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = 0.0;
    const double* xrow = base + (size_t)g * stride + t;   // X[c = g][. + t]
    for (int k = 0; k < np; ++k) {
        const int b = k & 1;
        mbar_wait(mbar + b, (phase >> b) & 1u);
        phase ^= 1u << b;
        const int j0 = k * COOP_PANEL_COLS;
        const int ncols = (d - j0 < COOP_PANEL_COLS) ? d - j0 : COOP_PANEL_COLS;   // multiple of 4
        if (active) {
            const double* __restrict__ pan = panels + (size_t)b * pan_elems + (size_t)t * pstride + i0 + g;   // A[j0 + t][i0 + g]
#pragma unroll 2
            for (int kk = 0; kk < ncols; kk += 4) {
                const double a = xrow[j0 + kk];
                const double* __restrict__ pk = pan + (size_t)kk * pstride;
                const double b0 = pk[0], b1 = pk[8], b2 = pk[16], b3 = pk[24];
                coop_dmma_m8n8k4(acc[0][0], acc[0][1], a, b0);
                coop_dmma_m8n8k4(acc[1][0], acc[1][1], a, b1);
                coop_dmma_m8n8k4(acc[2][0], acc[2][1], a, b2);
                coop_dmma_m8n8k4(acc[3][0], acc[3][1], a, b3);
            }
        }
        coop_barrier<NW>();   // every warp is done with buffer b
        if (warp == 0 && k + 2 < np) {
            const int j2 = (k + 2) * COOP_PANEL_COLS;
            const int nc2 = (d - j2 < COOP_PANEL_COLS) ? d - j2 : COOP_PANEL_COLS;
            bulk_load_panel_cols(panels + (size_t)b * pan_elems, A + (size_t)j2 * d, nc2, d, pstride, mbar + b, lane);
        }
    }
    if (active) {   // accumulator fragment: chain g, rows i0 + 8n + 2t, + 1
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int i = i0 + 8 * n + 2 * t;
            if (i < d) base[(size_t)g * stride + half + i] = acc[n][0];
            if (i + 1 < d) base[(size_t)g * stride + half + i + 1] = acc[n][1];
        }
    }
}

}  // namespace mcmcb200
