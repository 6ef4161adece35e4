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
    int lane;
    int d;
    double* scr;  // per-warp shared scratch (>= d doubles) for target functors that need all of x
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

}  // namespace mcmcb200
