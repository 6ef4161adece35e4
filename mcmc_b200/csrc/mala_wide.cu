// Chain-batched MALA for dense quadratic targets (linreg / dense_gauss), any n_dim up to 2048, M = I.
//
// The register-resident kernel (mala.cu) gives every chain its own warp and lets that warp stream the target's
// d x d matrix from L2 for each gradient: fine for a few chains, but at BASELINE config 3 (d = 1024, 16384 chains)
// that is 16384 x 8 MB of L2 traffic per draw.  All chains multiply the SAME matrix, so this path keeps the chains
// in lock-step and evaluates the gradient of every chain at once as one fp64 GEMM on the tensor cores:
//
//     AY[C x d] = Y[C x d] * A[d x d]          (A symmetric: the regression's X'X/s^2 + I/t^2, or a precision matrix)
//
// with hand-written DMMA (mma.sync.aligned.m8n8k4 f64) — tcgen05 has no f64 kind, so the legacy tensor path is the
// fp64 tensor path on sm_100a.  Per draw (src/mala.cpp:149-186 in the cancelled form of mala.cu / oracle.cpp):
//   1. gemm:  AY = Y A
//   2. rows:  one CTA per chain: log pi(y) = y.(b - AY/2), grad = b - AY, mu(y) = y + eps^2 grad / 2,
//             q1 = |x - mu(y)|^2 / eps^2, q2 = |y - mu(x)|^2 / eps^2, accept test, state update, draws_out row,
//             then the NEXT proposal y' = mu(x) + eps z (Philox or tape) — so a draw costs two launches.
// State (X, MU(X), Y, AY) lives chain-major in a workspace in HBM; log pi(x) and the accept counters are per-chain
// scalars.  Arithmetic is FMA-contracted (the GEMM accumulates in tensor-core order), so this path is held to the
// 1e-10 contract tolerance, not to bit-exactness.
#include "engine.h"
#include "rng.cuh"
#include "dgemm.h"
#include <math_constants.h>
#include <cstdio>
#include <cstdlib>

namespace mcmcb200
{

// ------------------------------------------------------------------------------------------------ DMMA GEMM
constexpr int GM = 128, GN = 64, GK = 16;       // CTA tile
constexpr int G_THREADS = 256;                   // 8 warps: 4 (M) x 2 (N), warp tile 32 x 32
constexpr int G_ASTR = GK + 4;                   // padded strides (doubles): conflict-free 64-bit fragment loads
constexpr int G_BSTR = GN + 4;
constexpr int G_STAGE_DOUBLES = GM * G_ASTR + GK * G_BSTR;
constexpr int G_STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C[M x N] = Y[M x K] * A[K x N], all row-major, leading dimensions = K for Y and N for A, C (here N == K == d).
// Requires d even and 16-byte aligned bases (checked by the launcher).
__global__ void __launch_bounds__(G_THREADS, 2) dgemm_dmma_kernel(const double* __restrict__ Y, const double* __restrict__ Amat,
                                                                  double* __restrict__ Cout, int M, int d)
{
    extern __shared__ __align__(16) double gsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;          // warp tile origin (wm*32, wn*32)
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int nk = (d + GK - 1) / GK;

    auto load_stage = [&](int stage, int kt) {
        double* As = gsm + (size_t)stage * G_STAGE_DOUBLES;
        double* Bs = As + GM * G_ASTR;
        const int k0 = kt * GK;
        // A tile: GM rows x GK doubles = GM x 8 chunks of 16 B
#pragma unroll
        for (int c = tid; c < GM * (GK / 2); c += G_THREADS) {
            const int r = c / (GK / 2), cc = (c % (GK / 2)) * 2;
            double* dst = As + r * G_ASTR + cc;
            const int gr = m0 + r, gc = k0 + cc;
            if (gr < M && gc + 1 < d) cp_async16(dst, Y + (size_t)gr * d + gc);
            else {
                dst[0] = (gr < M && gc < d) ? Y[(size_t)gr * d + gc] : 0.0;
                dst[1] = 0.0;
            }
        }
        // B tile: GK rows x GN doubles = GK x 32 chunks
#pragma unroll
        for (int c = tid; c < GK * (GN / 2); c += G_THREADS) {
            const int r = c / (GN / 2), cc = (c % (GN / 2)) * 2;
            double* dst = Bs + r * G_BSTR + cc;
            const int gr = k0 + r, gc = n0 + cc;
            if (gr < d && gc + 1 < d) cp_async16(dst, Amat + (size_t)gr * d + gc);
            else {
                dst[0] = (gr < d && gc < d) ? Amat[(size_t)gr * d + gc] : 0.0;
                dst[1] = 0.0;
            }
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < G_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();
        {   // prefetch tile kt + STAGES - 1 into the slot freed by iteration kt - 1
            const int nxt = kt + G_STAGES - 1;
            if (nxt < nk) load_stage(nxt % G_STAGES, nxt);
            cp_async_commit();
        }
        const double* As = gsm + (size_t)(kt % G_STAGES) * G_STAGE_DOUBLES;
        const double* Bs = As + GM * G_ASTR;
#pragma unroll
        for (int kk = 0; kk < GK / 4; ++kk) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[(wm * 32 + i * 8 + g) * G_ASTR + kk * 4 + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[(kk * 4 + t) * G_BSTR + wn * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + wm * 32 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn * 32 + j * 8 + 2 * t;
            if (row < M && col + 1 < d) *reinterpret_cast<double2*>(Cout + (size_t)row * d + col) = make_double2(acc[i][j][0], acc[i][j][1]);
            else if (row < M && col < d) Cout[(size_t)row * d + col] = acc[i][j][0];
        }
    }
}

// ------------------------------------------------------------------------------------------------ row kernels
constexpr int R_THREADS = 128;  // one CTA per chain

struct WideArgs {
    long long n_chains;
    int d;
    long long chain_offset;
    RngArgs rng;
    const double* bvec;   // b (linreg) or null (dense_gauss)
    double eps;
    double* X;            // [C][d] current state
    double* MX;           // [C][d] mu(x)
    double* Y;            // [C][d] proposal
    const double* AY;     // [C][d] A y (or A x during initialisation)
    double* LP;           // [C] log pi(x)
    long long* n_accept;  // [C]
    double* draws;        // [C][n_keep][d]
    double* logp;         // [C][n_keep] or null
    long long n_keep, n_burnin;
    const double* x0;
    int broadcast_x0;
};

// sum over the CTA in a fixed order (warp butterflies, then warps 0..3)
__device__ __forceinline__ double block_sum(double v, double* red)
{
    v = warp_sum<false>(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return ((red[0] + red[1]) + red[2]) + red[3];
}

// normals of draw t for this chain into z (thread-striped pairs), and uniform #0 (valid after the call on all threads)
template <int EPT, int RNGM>
__device__ __forceinline__ void row_normals(const WideArgs& a, long long chain, long long t, const double2* __restrict__ tab, double (&z)[EPT],
                                            unsigned* spare_sm, long long tape_pos)
{
    const int d = a.d, tid = threadIdx.x;
    if (RNGM == RNG_PHILOX) {
        const unsigned gchain = (unsigned)(a.chain_offset + chain);
#pragma unroll
        for (int m0 = 0; m0 < EPT / 2; m0 += 2) {
            BmPair b[2];
            double z0[2], z1[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int q = (m0 + i) * R_THREADS + tid;
                unsigned r[4];
                philox4x32_10((unsigned)q, (unsigned)(t + 1), gchain, 0u, a.rng, r);
                if (m0 + i == 0 && tid < 2) spare_sm[tid] = ((r[1] & 0xfffu) << 12) | (r[3] & 0xfffu);
                b[i].setup(r, tab);
            }
            bm_eval<2>(b, z0, z1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int q = (m0 + i) * R_THREADS + tid;
                if (m0 + i < EPT / 2) {
                    z[2 * (m0 + i)] = (2 * q < d) ? z0[i] : 0.0;
                    z[2 * (m0 + i) + 1] = (2 * q + 1 < d) ? z1[i] : 0.0;
                }
            }
        }
    } else {
        const double* tp = a.rng.tape + chain * a.rng.tape_stride + tape_pos;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int j = (k >> 1) * 2 * R_THREADS + 2 * tid + (k & 1);
            z[k] = (j < d) ? tp[j] : 0.0;
        }
    }
}

template <int EPT> __device__ __forceinline__ void row_load(const double* __restrict__ src, int d, double (&v)[EPT])
{
#pragma unroll
    for (int m = 0; m < EPT / 2; ++m) {
        const int j = m * 2 * R_THREADS + 2 * threadIdx.x;
        if (j + 1 < d) {
            const double2 u = *reinterpret_cast<const double2*>(src + j);
            v[2 * m] = u.x;
            v[2 * m + 1] = u.y;
        } else {
            v[2 * m] = (j < d) ? src[j] : 0.0;
            v[2 * m + 1] = 0.0;
        }
    }
}
template <int EPT> __device__ __forceinline__ void row_store(double* __restrict__ dst, int d, const double (&v)[EPT])
{
#pragma unroll
    for (int m = 0; m < EPT / 2; ++m) {
        const int j = m * 2 * R_THREADS + 2 * threadIdx.x;
        if (j + 1 < d) *reinterpret_cast<double2*>(dst + j) = make_double2(v[2 * m], v[2 * m + 1]);
        else if (j < d) dst[j] = v[2 * m];
    }
}

// phase 0 (t = -1): AY holds A x0.  Initialise LP, MX and emit the first proposal.
// phase 1 (t >= 0): AY holds A y_t.  Accept test for draw t, outputs, next proposal for draw t+1.
template <int EPT, int RNGM, bool INIT>
__global__ void __launch_bounds__(R_THREADS) mala_rows_kernel(const __grid_constant__ WideArgs a, const long long t)
{
    __shared__ double2 rng_tab[RNGM == RNG_PHILOX ? RNG_TAB_DOUBLE2 : 1];
    __shared__ double red[4];
    __shared__ unsigned spare_sm[2];
    if (RNGM == RNG_PHILOX) {
        build_rng_tables(rng_tab);
        __syncthreads();
    }
    const long long chain = blockIdx.x;
    const int d = a.d;
    const size_t row = (size_t)chain * d;
    const double eps = a.eps, e2 = eps * eps, he2 = 0.5 * e2;
    double x[EPT], mx[EPT], ay[EPT], bb[EPT];
    if (a.bvec) row_load<EPT>(a.bvec, d, bb);
    else {
#pragma unroll
        for (int k = 0; k < EPT; ++k) bb[k] = 0.0;
    }
    row_load<EPT>(a.AY + row, d, ay);
    const long long tape_stride_per_draw = d + 1;

    if (INIT) {
        row_load<EPT>(a.x0 + (a.broadcast_x0 ? 0 : row), d, x);
        double lp = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            lp = fma(x[k], bb[k] - 0.5 * ay[k], lp);             // log pi = x.(b - A x / 2)
            mx[k] = fma(he2, bb[k] - ay[k], x[k]);               // mu(x) = x + eps^2 (b - A x)/2
        }
        lp = block_sum(lp, red);
        if (threadIdx.x == 0) { a.LP[chain] = lp; a.n_accept[chain] = 0; }
        row_store<EPT>(a.X + row, d, x);
        row_store<EPT>(a.MX + row, d, mx);
    } else {
        double y[EPT], my[EPT];
        row_load<EPT>(a.X + row, d, x);
        row_load<EPT>(a.MX + row, d, mx);
        row_load<EPT>(a.Y + row, d, y);
        double part = 0.0, lp1 = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            lp1 = fma(y[k], bb[k] - 0.5 * ay[k], lp1);
            my[k] = fma(he2, bb[k] - ay[k], y[k]);
            const double r1 = x[k] - my[k], r2 = y[k] - mx[k];
            part = fma(r1, r1, part);                            // q1 - q2 numerator
            part = fma(-r2, r2, part);
        }
        const double LP = a.LP[chain];
        // one block reduction of (LP1 - LP) - (q1 - q2)/(2 eps^2); LP is added once (by thread 0's partial)
        double dl = block_sum(lp1 - 0.5 * part / e2, red);
        const double lp1_tot = block_sum(lp1, red);
        dl -= LP;
        // uniform #0 of draw t: the spare bits of this draw's normal blocks 0 and 1 (Philox) or the tape entry after the normals
        double u;
        if (RNGM == RNG_PHILOX) {
            const unsigned gchain = (unsigned)(a.chain_offset + chain);
            unsigned r0[4], r1[4];
            philox4x32_10(0u, (unsigned)(t + 1), gchain, 0u, a.rng, r0);
            philox4x32_10(1u, (unsigned)(t + 1), gchain, 0u, a.rng, r1);
            const unsigned s0 = ((r0[1] & 0xfffu) << 12) | (r0[3] & 0xfffu), s1 = ((r1[1] & 0xfffu) << 12) | (r1[3] & 0xfffu);
            const double sd = __hiloint2double(0x43300000 | (s0 >> 8), (s0 << 24) | s1) - 4503599627370496.0;
            u = fma(sd, 3.5527136788005009e-15, 1.7763568394002505e-15);
        } else {
            u = a.rng.tape[chain * a.rng.tape_stride + t * tape_stride_per_draw + d];
        }
        bool acc = false;   // a non-finite proposal density rejects (src/mala.cpp:164-166); dl = +inf accepts (same rule as mala.cu)
        if (isfinite(lp1_tot)) {
            acc = u < 1.0 + dl;
            if (!acc) acc = (fabs(dl) <= 1.7976931348623157e308) && (u < exp(dl));
        }
        if (acc) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) { x[k] = y[k]; mx[k] = my[k]; }
            row_store<EPT>(a.X + row, d, x);
            row_store<EPT>(a.MX + row, d, mx);
            if (threadIdx.x == 0) a.LP[chain] = lp1_tot;
        }
        if (t >= a.n_burnin) {
            const long long r = t - a.n_burnin;
            row_store<EPT>(a.draws + ((size_t)chain * a.n_keep + r) * d, d, x);
            if (threadIdx.x == 0) {
                if (a.logp) a.logp[chain * a.n_keep + r] = acc ? lp1_tot : LP;
                if (acc) a.n_accept[chain] += 1;
            }
        }
    }
    // next proposal: y = mu(x) + eps z_{t+1}
    double z[EPT];
    row_normals<EPT, RNGM>(a, chain, t + 1, rng_tab, z, spare_sm, (t + 1) * tape_stride_per_draw);
#pragma unroll
    for (int k = 0; k < EPT; ++k) z[k] = fma(eps, z[k], mx[k]);
    row_store<EPT>(a.Y + row, d, z);
}

template <int EPT, int RNGM> static int launch_rows(const WideArgs& a, long long t, bool init, cudaStream_t st)
{
    if (init) mala_rows_kernel<EPT, RNGM, true><<<(unsigned)a.n_chains, R_THREADS, 0, st>>>(a, t);
    else mala_rows_kernel<EPT, RNGM, false><<<(unsigned)a.n_chains, R_THREADS, 0, st>>>(a, t);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <int RNGM> static int launch_rows_ept(const WideArgs& a, long long t, bool init, cudaStream_t st)
{
    const int ept = 2 * ((a.d + 2 * R_THREADS - 1) / (2 * R_THREADS));
    switch (ept) {
    case 2: return launch_rows<2, RNGM>(a, t, init, st);
    case 4: return launch_rows<4, RNGM>(a, t, init, st);
    case 6: case 8: return launch_rows<8, RNGM>(a, t, init, st);
    default: return launch_rows<16, RNGM>(a, t, init, st);
    }
}

// C[M x d] = Y[M x d] * A[d x d] (row-major) on `stream`: the fp64 tensor-core GEMM above, for the other chain-batched paths
int launch_dgemm_dmma(const double* Y, const double* Amat, double* Cout, long long M, int d, cudaStream_t stream)
{
    if ((d & 1) || ((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(Amat) | reinterpret_cast<uintptr_t>(Cout)) & 15)) {
        set_error("dgemm: n_dim must be even and the operands 16-byte aligned");
        return MCMCB200_ERR_UNSUPPORTED;
    }
    const size_t gsmem = (size_t)G_STAGES * G_STAGE_DOUBLES * sizeof(double);
    static bool attr_set[64] = {};   // per device, once: the call may sit inside a stream capture
    int dev = 0;
    MCMCB200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        MCMCB200_CUDA_TRY(cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const dim3 ggrid((d + GN - 1) / GN, (unsigned)((M + GM - 1) / GM));
    dgemm_dmma_kernel<<<ggrid, G_THREADS, gsmem, stream>>>(Y, Amat, Cout, (int)M, d);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

long long mala_wide_work_doubles(long long n_chains, int d) { return 4 * n_chains * (long long)d + n_chains; }

bool mala_wide_supported(int target_id, int d, bool has_precond)
{
    return !has_precond && (target_id == MCMCB200_TARGET_LINREG || target_id == MCMCB200_TARGET_DENSE_GAUSS) && d >= 2 && d <= 2048 &&
           (d % 2 == 0);
}

// Runs the whole chain-batched MALA job; returns the number of kernel launches through *launches.
int launch_mala_wide(const MalaLaunch& m, double* work, int* launches)
{
    const int d = m.d;
    const long long C = m.n_chains;
    WideArgs a;
    a.n_chains = C; a.d = d; a.chain_offset = m.chain_offset; a.rng = m.rng;
    a.bvec = (m.target_id == MCMCB200_TARGET_LINREG) ? m.tdata + (size_t)d * d : nullptr;
    a.eps = m.eps;
    a.X = work; a.MX = work + (size_t)C * d; a.Y = work + 2 * (size_t)C * d;
    double* AY = work + 3 * (size_t)C * d;
    a.AY = AY; a.LP = work + 4 * (size_t)C * d;
    a.n_accept = m.n_accept; a.draws = m.draws; a.logp = m.logp; a.n_keep = m.n_keep; a.n_burnin = m.n_burnin;
    a.x0 = m.x0; a.broadcast_x0 = m.broadcast_x0;
    if (m.broadcast_x0) { set_error("mala (chain-batched path): broadcast_initial is not supported"); return MCMCB200_ERR_UNSUPPORTED; }

    const size_t gsmem = (size_t)G_STAGES * G_STAGE_DOUBLES * sizeof(double);
    MCMCB200_CUDA_TRY(cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
    const dim3 ggrid((d + GN - 1) / GN, (unsigned)((C + GM - 1) / GM));
    if (getenv("MCMCB200_DEBUG")) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, dgemm_dmma_kernel);
        fprintf(stderr, "[mcmc_b200] dgemm attrs: %s regs=%d maxDyn=%d static=%zu maxThreads=%d bin=%d grid=(%u,%u) smem=%zu C=%lld d=%d x0=%p tdata=%p work=%p\n",
                cudaGetErrorString(e), fa.numRegs, fa.maxDynamicSharedSizeBytes, fa.sharedSizeBytes, fa.maxThreadsPerBlock, fa.binaryVersion,
                ggrid.x, ggrid.y, gsmem, C, d, (const void*)m.x0, (const void*)m.tdata, (void*)work);
    }
    int nl = 0;
    auto gemm = [&](const double* Yin) -> int {
        dgemm_dmma_kernel<<<ggrid, G_THREADS, gsmem, m.stream>>>(Yin, m.tdata, AY, (int)C, d);
        MCMCB200_CUDA_TRY(cudaGetLastError());
        ++nl;
        return MCMCB200_OK;
    };
    auto rows = [&](long long t, bool init) -> int {
        ++nl;
        return (m.rng.mode == RNG_PHILOX) ? launch_rows_ept<RNG_PHILOX>(a, t, init, m.stream) : launch_rows_ept<RNG_TAPE>(a, t, init, m.stream);
    };
    int rc;
    if ((rc = gemm(m.x0))) return rc;          // A x0
    if ((rc = rows(-1, true))) return rc;      // LP(x0), mu(x0), first proposal
    const long long n_total = m.n_burnin + m.n_keep;
    for (long long t = 0; t < n_total; ++t) {
        if ((rc = gemm(a.Y))) return rc;
        if ((rc = rows(t, false))) return rc;
    }
    *launches = nl;
    return MCMCB200_OK;
}

}  // namespace mcmcb200
