// Warp-level building blocks shared by the sampler kernels (sm_100a).
//
// Layout contract ("one warp per chain", BASELINE.json north_star): a chain's d-vector
// is striped over the 32 lanes of its warp in 128-bit lane pairs — element j lives on
// lane (j % 64) / 2 in register slot k = 2*(j / 64) + (j % 2), so slot pair (2m, 2m+1)
// of lane l holds elements (64m + 2l, 64m + 2l + 1) and one double2 access per lane
// moves a 512-byte contiguous segment per warp instruction.  EPL (elements per lane)
// = 2*ceil(d/64); slots with j >= d hold 0.0 and never change.
//
// Reductions: each lane adds its slots in increasing k, then a 5-stage xor butterfly
// (offsets 16,8,4,2,1) — the order oracle/host_targets.hpp::reduce_sum(SUM_WARP) restates.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mcmcb200
{

constexpr unsigned FULL = 0xffffffffu;

// Arithmetic policy.  STRICT reproduces the reference's un-contracted IEEE operation
// order (oracle built with -ffp-contract=off); FAST lets every a*b+c fuse.
template <bool STRICT> struct Ar {
    static __device__ __forceinline__ double mul(double a, double b) { return STRICT ? __dmul_rn(a, b) : a * b; }
    static __device__ __forceinline__ double add(double a, double b) { return STRICT ? __dadd_rn(a, b) : a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return STRICT ? __dsub_rn(a, b) : a - b; }
    // a*b + c
    static __device__ __forceinline__ double mad(double a, double b, double c)
    {
        return STRICT ? __dadd_rn(__dmul_rn(a, b), c) : fma(a, b, c);
    }
};

struct WarpCtx {
    static constexpr int coop = 0;
    int lane;
    int d;
    double* scr;  // per-warp shared scratch (>= d doubles) for target functors that need all of x
};
// Context of a kernel whose NW warps per CTA evaluate dense products cooperatively (coop_gemv below): warp c stages its
// vector at coop_base + c*coop_stride (== its scr) and receives the product at + coop_stride/2.  panels / mbar: the two
// TMA panel buffers (COOP_PANEL_COLS columns of the matrix each) and their mbarriers, or null for the direct-load path.
template <int NW> struct CoopWarpCtx : WarpCtx {
    static constexpr int coop = NW;
    int warp;
    double* coop_base;
    int coop_stride;
    double* panels;
    unsigned long long* mbar;
    int panel_stride;         // doubles between consecutive matrix columns in a panel buffer (d, or d + 4 on the DMMA path)
    int dmma;                 // FAST arithmetic: the NW-chain product runs on the fp64 tensor cores (coop_gemv_body_dmma)
    const int* n_active;      // chains of this CTA still running (shared memory)
    int* want;                // products requested and not yet served (shared memory): busy warps poll it, see coop_round
    int* round_seq;           // number of rounds this CTA has run (shared memory): a request posted at sequence s is served once it is > s
    mutable int prefetched;   // the product for the staged vector was requested earlier and has been served: dense_matvec only reads it
    mutable unsigned phase;   // bit b = parity of panel buffer b's next completed phase (identical in every thread)
};

__device__ __forceinline__ int elem_index(int lane, int k) { return (k >> 1) * 64 + 2 * lane + (k & 1); }

template <bool STRICT> __device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = Ar<STRICT>::add(v, __shfl_xor_sync(FULL, v, off));
    return v;
}

// two independent butterflies interleaved (ILP)
template <bool STRICT> __device__ __forceinline__ void warp_sum2(double& a, double& b)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ta = __shfl_xor_sync(FULL, a, off);
        const double tb = __shfl_xor_sync(FULL, b, off);
        a = Ar<STRICT>::add(a, ta);
        b = Ar<STRICT>::add(b, tb);
    }
}

// sum_j a_j * b_j in the layout order
template <int EPL, bool STRICT> __device__ __forceinline__ double lane_dot(const double (&a)[EPL], const double (&b)[EPL])
{
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < EPL; ++k) s = Ar<STRICT>::mad(a[k], b[k], s);
    return s;
}
template <int EPL, bool STRICT> __device__ __forceinline__ double warp_dot(const double (&a)[EPL], const double (&b)[EPL])
{
    return warp_sum<STRICT>(lane_dot<EPL, STRICT>(a, b));
}

// ---- global <-> register striping ----------------------------------------------------

template <int EPL> __device__ __forceinline__ void load_vec(const double* __restrict__ src, int d, int lane, double (&x)[EPL])
{
    const bool vec_ok = ((d & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m) {
        const int j = m * 64 + 2 * lane;
        if (vec_ok) {
            if (j < d) {
                const double2 v = *reinterpret_cast<const double2*>(src + j);
                x[2 * m] = v.x;
                x[2 * m + 1] = v.y;
            } else {
                x[2 * m] = 0.0;
                x[2 * m + 1] = 0.0;
            }
        } else {
            x[2 * m] = (j < d) ? src[j] : 0.0;
            x[2 * m + 1] = (j + 1 < d) ? src[j + 1] : 0.0;
        }
    }
}

// same striping for memory the kernel also WRITES (population matrices, work areas): no __restrict__ / read-only promise,
// L2-coherent loads (ld.global.cg), so a row stored by the warp and re-read after __syncwarp() is never served stale
template <int EPL> __device__ __forceinline__ void load_vec_rw(const double* src, int d, int lane, double (&x)[EPL])
{
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m) {
        const int j = m * 64 + 2 * lane;
        x[2 * m] = (j < d) ? __ldcg(src + j) : 0.0;
        x[2 * m + 1] = (j + 1 < d) ? __ldcg(src + j + 1) : 0.0;
    }
}

// full-tile variants: d == 32*EPL and 16-byte aligned rows (checked on the host), no predicates
template <int EPL> __device__ __forceinline__ void load_vec_full(const double* __restrict__ src, int lane, double (&x)[EPL])
{
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m) {
        const double2 v = *reinterpret_cast<const double2*>(src + m * 64 + 2 * lane);
        x[2 * m] = v.x;
        x[2 * m + 1] = v.y;
    }
}
template <int EPL> __device__ __forceinline__ void store_vec_full(double* __restrict__ dst, int lane, const double (&x)[EPL])
{
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m) *reinterpret_cast<double2*>(dst + m * 64 + 2 * lane) = make_double2(x[2 * m], x[2 * m + 1]);
}

template <int EPL> __device__ __forceinline__ void store_vec(double* __restrict__ dst, int d, int lane, const double (&x)[EPL])
{
    const bool vec_ok = ((d & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m) {
        const int j = m * 64 + 2 * lane;
        if (vec_ok) {
            if (j < d) *reinterpret_cast<double2*>(dst + j) = make_double2(x[2 * m], x[2 * m + 1]);
        } else {
            if (j < d) dst[j] = x[2 * m];
            if (j + 1 < d) dst[j + 1] = x[2 * m + 1];
        }
    }
}

// registers -> per-warp shared vector (index = element index)
template <int EPL> __device__ __forceinline__ void stage_vec(double* __restrict__ smem, int d, int lane, const double (&x)[EPL])
{
    __syncwarp();
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m) {
        const int j = m * 64 + 2 * lane;
        if (j < d) smem[j] = x[2 * m];
        if (j + 1 < d) smem[j + 1] = x[2 * m + 1];
    }
    __syncwarp();
}

// y_i = sum_j A[j*d + i] * (alpha * v_j), j increasing: column-major operator applied to the
// staged vector v (the order of gemv_scaled / gemv_plain in oracle/oracle.cpp).  A symmetric
// row-major matrix is its own column-major image.
template <int EPL, bool STRICT>
__device__ __forceinline__ void gemv_cm(const double* __restrict__ A, int d, int lane, const double* __restrict__ v_smem,
                                        double alpha, double (&y)[EPL])
{
#pragma unroll
    for (int k = 0; k < EPL; ++k) y[k] = 0.0;
    const bool vec_ok = ((d & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    for (int j = 0; j < d; ++j) {
        const double t = Ar<STRICT>::mul(alpha, v_smem[j]);
        const double* __restrict__ col = A + (size_t)j * (size_t)d;
#pragma unroll
        for (int m = 0; m < EPL / 2; ++m) {
            const int i = m * 64 + 2 * lane;
            double a0 = 0.0, a1 = 0.0;
            if (vec_ok) {
                if (i < d) {
                    const double2 a = __ldg(reinterpret_cast<const double2*>(col + i));
                    a0 = a.x;
                    a1 = a.y;
                }
            } else {
                if (i < d) a0 = __ldg(col + i);
                if (i + 1 < d) a1 = __ldg(col + i + 1);
            }
            y[2 * m] = Ar<STRICT>::mad(a0, t, y[2 * m]);
            y[2 * m + 1] = Ar<STRICT>::mad(a1, t, y[2 * m + 1]);
        }
    }
    // padding slots stay exactly 0 even if v carried inf/nan
#pragma unroll
    for (int k = 0; k < EPL; ++k)
        if (elem_index(lane, k) >= d) y[k] = 0.0;
}

// y_i = sum_j ((rs_i * A[j*d + i]) * post) * v_j, j increasing: what the reference's  (scalar * diag) * Matrix * vector
// chains evaluate to with box constraints (bounded MALA with a dense precond_mat: drift ((eps^2 J) M) grad, noise
// ((eps chol J) sqrtM) z — the diagonal factor scales the ROWS of the dense matrix before the product; oracle.cpp mala_mean).
template <int EPL, bool STRICT>
__device__ __forceinline__ void gemv_cm_rowscaled(const double* __restrict__ A, int d, int lane, const double* __restrict__ v_smem,
                                                  const double (&rs)[EPL], double post, double (&y)[EPL])
{
#pragma unroll
    for (int k = 0; k < EPL; ++k) y[k] = 0.0;
    for (int j = 0; j < d; ++j) {
        const double t = v_smem[j];
        const double* __restrict__ col = A + (size_t)j * (size_t)d;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const int i = elem_index(lane, k);
            const double a = (i < d) ? __ldg(col + i) : 0.0;
            y[k] = Ar<STRICT>::mad(Ar<STRICT>::mul(Ar<STRICT>::mul(rs[k], a), post), t, y[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < EPL; ++k)
        if (elem_index(lane, k) >= d) y[k] = 0.0;
}


// ---- CTA-cooperative dense product -----------------------------------------------------------------------------
// y_c = A x_c for the NW chains of a CTA at once (A column-major d x d, read ONCE per call for all chains instead of
// once per chain): every warp of the CTA calls coop_gemv the same number of times; between the two named barriers
// warp w computes rows {lane + 32 (w + NW m)} of all NW products from the vectors staged in shared memory.  Each
// output element is accumulated over j = 0..d-1 in increasing order with one multiply-add per term — exactly the order
// of gemv_cm — so the result is bit-identical to the per-warp product in both arithmetic modes.
// The matrix is streamed through shared memory in panels of COOP_PANEL_COLS columns (one contiguous d*32*8-byte block
// of the column-major array) by TMA bulk copies (cp.async.bulk + mbarrier complete_tx), double-buffered: with only
// NW = 8 resident warps per SM, per-lane global loads cannot keep enough bytes in flight to cover the L2 latency
// (measured: 38 us per product at d = 256 with direct loads), a 64 KB bulk copy per panel can.
// Barrier 1 is used (bar.sync 1, 32*NW): all 32*NW threads of the CTA must arrive, from any call site.
constexpr int COOP_PANEL_COLS = 32;

template <int NW> __device__ __forceinline__ void coop_barrier() { asm volatile("bar.sync 1, %0;" : : "n"(32 * NW) : "memory"); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" : : "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" : : : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}"
        : : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one thread: announce `bytes` on the mbarrier and start the bulk copy global -> shared that will complete them
__device__ __forceinline__ void bulk_load_panel(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" : : "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 : : "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// direct-load body (ragged / odd n_dim, or no panel buffers): four columns in flight per lane
template <int NW, bool STRICT>
__device__ __forceinline__ void coop_gemv_body(const double* __restrict__ A, int d, int warp, int lane, double* __restrict__ base, int stride)
{
    const int half = stride >> 1;
    for (int i0 = 32 * warp; i0 < d; i0 += 32 * NW) {
        const int i = i0 + lane;
        const bool row_ok = i < d;
        double acc[NW];
#pragma unroll
        for (int c = 0; c < NW; ++c) acc[c] = 0.0;
        const double* __restrict__ col = A + (row_ok ? i : 0);
        int j = 0;
        for (; j + 4 <= d; j += 4) {
            double a4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a4[u] = __ldg(col + (size_t)(j + u) * (size_t)d);
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
#pragma unroll
                for (int c = 0; c < NW; ++c) {
                    const double2 xv = *reinterpret_cast<const double2*>(base + (size_t)c * stride + j + u);
                    acc[c] = Ar<STRICT>::mad(a4[u], xv.x, acc[c]);
                    acc[c] = Ar<STRICT>::mad(a4[u + 1], xv.y, acc[c]);
                }
            }
        }
        for (; j < d; ++j) {
            const double aj = __ldg(col + (size_t)j * (size_t)d);
#pragma unroll
            for (int c = 0; c < NW; ++c) acc[c] = Ar<STRICT>::mad(aj, base[(size_t)c * stride + j], acc[c]);
        }
        if (row_ok) {
#pragma unroll
            for (int c = 0; c < NW; ++c) base[(size_t)c * stride + half + i] = acc[c];
        }
    }
}

// TMA-panel body: requires d even, d <= 32*NW (one row per lane), A 16-byte aligned.  Called by all 32*NW threads
// after every chain's x is staged; thread 0 is the producer.  Ends with every panel buffer free again.
template <int NW, bool STRICT>
__device__ __forceinline__ void coop_gemv_body_tma(const double* __restrict__ A, int d, int warp, int lane, double* __restrict__ base,
                                                   int stride, double* __restrict__ panels, unsigned long long* mbar, unsigned& phase)
{
    const int half = stride >> 1;
    const int np = (d + COOP_PANEL_COLS - 1) / COOP_PANEL_COLS;
    const size_t pan_elems = (size_t)COOP_PANEL_COLS * d;
    const int i = 32 * warp + lane;
    const bool row_ok = i < d;
    const int ir = row_ok ? i : 0;
    double acc[NW];
#pragma unroll
    for (int c = 0; c < NW; ++c) acc[c] = 0.0;
    for (int k = 0; k < np; ++k) {
        const int b = k & 1;
        mbar_wait(mbar + b, (phase >> b) & 1u);
        phase ^= 1u << b;
        const double* __restrict__ pan = panels + (size_t)b * pan_elems + ir;
        const int j0 = k * COOP_PANEL_COLS;
        const int ncols = (d - j0 < COOP_PANEL_COLS) ? d - j0 : COOP_PANEL_COLS;   // even
#pragma unroll 4
        for (int jl = 0; jl < ncols; jl += 2) {
            const double a0 = pan[(size_t)jl * d], a1 = pan[(size_t)(jl + 1) * d];
#pragma unroll
            for (int c = 0; c < NW; ++c) {
                const double2 xv = *reinterpret_cast<const double2*>(base + (size_t)c * stride + j0 + jl);
                acc[c] = Ar<STRICT>::mad(a0, xv.x, acc[c]);
                acc[c] = Ar<STRICT>::mad(a1, xv.y, acc[c]);
            }
        }
        coop_barrier<NW>();   // every warp is done with buffer b
        if (threadIdx.x == 0 && k + 2 < np) {
            const int j2 = (k + 2) * COOP_PANEL_COLS;
            const int nc2 = (d - j2 < COOP_PANEL_COLS) ? d - j2 : COOP_PANEL_COLS;
            bulk_load_panel(panels + (size_t)b * pan_elems, A + (size_t)j2 * d, (unsigned)((size_t)nc2 * d * sizeof(double)), mbar + b);
        }
    }
    if (row_ok) {
#pragma unroll
        for (int c = 0; c < NW; ++c) base[(size_t)c * stride + half + i] = acc[c];
    }
}

// Defined in coop_dmma.cuh (included by the one kernel source that runs cooperative rounds, nuts.cu; every other translation
// unit only sees these declarations and never instantiates them): panel producer for dense / padded panel buffers and the
// fp64 tensor-core body of the round.
__device__ __forceinline__ void coop_panels_prologue(const double* __restrict__ A, int d, double* panels, int pstride, unsigned long long* mbar);
template <int NW>
__device__ __forceinline__ void coop_gemv_body_dmma(const double* __restrict__ A, int d, int warp, int lane, double* __restrict__ base, int stride,
                                                    double* __restrict__ panels, int pstride, unsigned long long* mbar, unsigned& phase);

// The whole cooperative round for one context; active chains (draining = false) and the drain loop of warps whose chain
// is finished (draining = true) call exactly this, so every warp of the CTA passes the same barriers.  *w.n_active (chains
// of the CTA still running) only changes between a round's last barrier and the next round's first one, so the value read
// after the first barrier is the same in every warp; it can only be zero when every warp is draining.
// A warp that needs a product posts it in *w.want before the first barrier; warps that are busy with work that needs
// no product (the NUTS tree replay) poll *w.want and attend the round as helpers (draining = true) within one step of
// their own work, so a requester never waits for another chain's whole replay.  The counter is cleared inside the
// round, when every warp is present and nobody can be posting.
template <bool STRICT, class Ctx> __device__ __forceinline__ bool coop_round(const double* __restrict__ A, const Ctx& w, bool draining)
{
    constexpr int NW = Ctx::coop;
    if (!draining && w.lane == 0) atomicAdd(w.want, 1);
    coop_barrier<NW>();   // every chain's x is staged
    if (draining && *reinterpret_cast<const volatile int*>(w.n_active) == 0) return false;
    if (threadIdx.x == 0) {   // visible to everyone after the last barrier
        *reinterpret_cast<volatile int*>(w.want) = 0;
        *reinterpret_cast<volatile int*>(w.round_seq) = *reinterpret_cast<volatile int*>(w.round_seq) + 1;
    }
    if (w.panels) {
        coop_panels_prologue(A, w.d, w.panels, w.panel_stride, w.mbar);
        if (!STRICT && w.dmma) coop_gemv_body_dmma<NW>(A, w.d, w.warp, w.lane, w.coop_base, w.coop_stride, w.panels, w.panel_stride, w.mbar, w.phase);
        else coop_gemv_body_tma<NW, STRICT>(A, w.d, w.warp, w.lane, w.coop_base, w.coop_stride, w.panels, w.mbar, w.phase);
    } else {
        coop_gemv_body<NW, STRICT>(A, w.d, w.warp, w.lane, w.coop_base, w.coop_stride);
    }
    coop_barrier<NW>();   // every chain's y is complete
    return true;
}

// y = A x through the cooperative path when the context asks for it (NW is the CTA's warp count), else per warp.
template <int EPL, bool STRICT, class Ctx>
__device__ __forceinline__ void dense_matvec(const double* __restrict__ A, const Ctx& w, const double (&x)[EPL], double (&y)[EPL])
{
    if constexpr (Ctx::coop > 0) {
        if (!w.prefetched) {   // (else: x was staged and its product requested ahead of time, and a round has served it — nuts.cu)
            stage_vec<EPL>(w.scr, w.d, w.lane, x);
            coop_round<STRICT>(A, w, false);
        }
        load_vec<EPL>(w.scr + (w.coop_stride >> 1), w.d, w.lane, y);
    } else {
        stage_vec<EPL>(w.scr, w.d, w.lane, x);
        gemv_cm<EPL, STRICT>(A, w.d, w.lane, w.scr, 1.0, y);
    }
}

}  // namespace mcmcb200
