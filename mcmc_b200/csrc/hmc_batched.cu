// Chain-batched HMC: every dense d x d product of the trajectory runs as ONE fp64 tensor-core GEMM over all chains.
//
// BASELINE north_star: "tensor cores used only for the dense M^-1 p product ... where M is a full d x d matrix".  The
// reference multiplies by its (inverse / square-root) mass matrix per chain and per leapfrog step
// (/root/reference/src/hmc.cpp:57-59 one-time INV / CHOL, :158 p = sqrtM z, :160,:184 K = p.(M^-1 p)/2, :171 x += (eps M^-1) p)
// and, for a dense quadratic target, evaluates the gradient with another d x d product.  HMC has a fixed trajectory
// length, so all chains of a call are in lock-step and each of these products is the same matrix applied to every chain:
//     P  = Z  * sqrtM'       MP = P * M^-1       AX = X * A        (chain-major [C][d] operands, row-major GEMM)
// done by dgemm_dmma_kernel (mala_wide.cu: mma.sync.m8n8k4.f64 — tcgen05 has no f64 kind — 33 TFLOP/s fp64 on B200),
// instead of C warps each streaming the matrix from L2 (hmc.cu: 4096 chains x (L+3) x d^2 x 8 B of L2 traffic per draw).
// Between the GEMMs one row kernel per phase (one CTA per chain) does the element-wise work: momentum refresh (Philox or
// tape), kicks, drifts, energies, the Metropolis test, the draws_out row.  One draw = L + 3 mass GEMMs (+ L + 1 target
// GEMMs for a dense target) and as many row launches; the launch sequence of a draw is captured in a CUDA graph once and
// replayed n_burnin + n_keep times (the draw index lives in device memory).
// Scope: FAST arithmetic (the GEMM accumulates in tensor-core order: held to the 1e-10 contract, not to bits), no box
// constraints, n_dim even and <= 2048 — which also lifts the 512-element limit of the register-resident kernels for dense
// targets / dense mass.  Targets: iso_gauss, diag_gauss (element-wise gradient), dense_gauss, linreg (GEMM gradient).
#include "engine.h"
#include "rng.cuh"
#include "dgemm.h"
#include "hmc_batched.h"
#include <math_constants.h>
#include <cstdlib>

namespace mcmcb200
{

constexpr int HB_THREADS = 128;

struct HbArgs {
    long long n_chains;
    int d;
    int target_id;
    const double* tdata;
    long long chain_offset;
    RngArgs rng;
    double eps;
    int n_leap;
    int dense_mass, dense_target;
    const double* x0;
    double* X;    // [C][d] current state
    double* XT;   // [C][d] trajectory position
    double* P;    // [C][d] momentum
    double* Z;    // [C][d] normals (dense mass) — aliases P when M = I
    double* MP;   // [C][d] M^-1 p (dense mass) — aliases P when M = I
    double* AX;   // [C][d] A xt (dense target)
    double* U0;   // [C] -log pi(x)
    double* H0;   // [C] U0 + K0 of the running draw
    double* U1;   // [C] -log pi(xt) at the end of the trajectory
    double* Uu;   // [C] the draw's uniform
    long long* n_accept;
    double* draws;
    double* logp;
    long long n_keep, n_burnin;
    int* t_dev;   // draw index (device), advanced by the last phase of a draw
};

__device__ __forceinline__ double hb_block_sum(double v, double* red)
{
    v = warp_sum<false>(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return ((red[0] + red[1]) + red[2]) + red[3];
}

// gradient of log pi and this element pair's part of log pi, from x and (dense targets) the product A x
__device__ __forceinline__ void hb_grad(const HbArgs& a, int j, double x, double ax, double& g, double& lp)
{
    switch (a.target_id) {
    case MCMCB200_TARGET_ISO_GAUSS: g = -x; lp = -0.5 * x * x; break;
    case MCMCB200_TARGET_DIAG_GAUSS: { const double w = __ldg(a.tdata + j); g = -w * x; lp = -0.5 * w * x * x; break; }
    case MCMCB200_TARGET_DENSE_GAUSS: g = -ax; lp = -0.5 * x * ax; break;
    default: { const double b = __ldg(a.tdata + (size_t)a.d * a.d + j); g = b - ax; lp = x * (b - 0.5 * ax); break; }   // linreg
    }
}

// Phases (one CTA per chain, thread owns element pairs q = tid, tid + 128, ...):
//   0 INIT  : X = x0; U0 = -log pi(x0) (needs AX = x0 A for dense targets)
//   1 BEGIN : z ~ N(0, I) -> Z; XT = X; the draw's uniform
//   2 KICK0 : K0 = p.MP/2, H0 = U0 + K0; p += (eps/2) grad(xt)                       (AX = A xt valid for dense targets)
//   3 DRIFT : xt += eps MP                                                            (dense targets: gradient needs a GEMM first)
//   4 KICK  : p += c grad(xt), c = eps (inner steps) or eps/2 (last, which also stores U1 = -log pi(xt))
//   5 END   : K1 = p.MP/2; accept iff u < exp(min(0.01, H0 - U1 - K1)); X, U0, n_accept, draws_out row; ++t
template <int PH, int RNGM> __global__ void __launch_bounds__(HB_THREADS) hmc_batched_rows(const HbArgs a, int last_kick)
{
    __shared__ double2 rng_tab[(PH == 1 && RNGM == RNG_PHILOX) ? RNG_TAB_DOUBLE2 : 1];
    __shared__ double red[4];
    __shared__ unsigned spare_sm[2];
    const long long chain = blockIdx.x;
    const int d = a.d, tid = threadIdx.x;
    const size_t row = (size_t)chain * d;
    const int t = *a.t_dev;
    const int npair = d >> 1;
    if (PH == 0) {
        double lp = 0.0;
        for (int q = tid; q < npair; q += HB_THREADS) {
            const double2 x = *reinterpret_cast<const double2*>(a.x0 + row + 2 * q);
            *reinterpret_cast<double2*>(a.X + row + 2 * q) = x;
            double2 ax = make_double2(0.0, 0.0);
            if (a.dense_target) ax = *reinterpret_cast<const double2*>(a.AX + row + 2 * q);
            double g, l0, l1;
            hb_grad(a, 2 * q, x.x, ax.x, g, l0);
            hb_grad(a, 2 * q + 1, x.y, ax.y, g, l1);
            lp += l0 + l1;
        }
        lp = hb_block_sum(lp, red);
        if (tid == 0) { a.U0[chain] = -lp; a.n_accept[chain] = 0; }
    } else if (PH == 1) {
        if (RNGM == RNG_PHILOX) {
            build_rng_tables(rng_tab);
            __syncthreads();
            const unsigned gchain = (unsigned)(a.chain_offset + chain);
            for (int q = tid; q < npair; q += HB_THREADS) {
                unsigned r[4];
                philox4x32_10((unsigned)q, (unsigned)(t + 1), gchain, 0u, a.rng, r);
                if (q < 2) spare_sm[q] = ((r[1] & 0xfffu) << 12) | (r[3] & 0xfffu);
                BmPair b[1];
                double z0[1], z1[1];
                b[0].setup(r, rng_tab);
                bm_eval<1>(b, z0, z1);
                *reinterpret_cast<double2*>(a.Z + row + 2 * q) = make_double2(z0[0], z1[0]);
                *reinterpret_cast<double2*>(a.XT + row + 2 * q) = *reinterpret_cast<const double2*>(a.X + row + 2 * q);
            }
            if (npair < 2 && tid == 1) {   // d = 2: block 1 carries no normals but its spare bits are part of the uniform
                unsigned r[4];
                philox4x32_10(1u, (unsigned)(t + 1), gchain, 0u, a.rng, r);
                spare_sm[1] = ((r[1] & 0xfffu) << 12) | (r[3] & 0xfffu);
            }
            __syncthreads();
            if (tid == 0) {
                const unsigned s0 = spare_sm[0], s1 = spare_sm[1];
                const double sd = __hiloint2double(0x43300000 | (s0 >> 8), (s0 << 24) | s1) - 4503599627370496.0;
                a.Uu[chain] = fma(sd, 3.5527136788005009e-15, 1.7763568394002505e-15);
            }
        } else {
            const double* tp = a.rng.tape + chain * a.rng.tape_stride + (long long)t * (d + 1);
            for (int q = tid; q < npair; q += HB_THREADS) {
                *reinterpret_cast<double2*>(a.Z + row + 2 * q) = make_double2(tp[2 * q], tp[2 * q + 1]);
                *reinterpret_cast<double2*>(a.XT + row + 2 * q) = *reinterpret_cast<const double2*>(a.X + row + 2 * q);
            }
            if (tid == 0) a.Uu[chain] = tp[d];
        }
    } else if (PH == 2 || PH == 4) {
        const double c = (PH == 2 || last_kick) ? 0.5 * a.eps : a.eps;
        const bool kick = a.n_leap > 0;   // L = 0: the proposal is the current state and the momentum is left alone (K1 = K0)
        double k0 = 0.0, lp = 0.0;
        for (int q = tid; q < npair; q += HB_THREADS) {
            double2 p = *reinterpret_cast<const double2*>(a.P + row + 2 * q);
            if (PH == 2) {
                const double2 mp = a.dense_mass ? *reinterpret_cast<const double2*>(a.MP + row + 2 * q) : p;
                k0 = fma(p.x, mp.x, fma(p.y, mp.y, k0));
            }
            const double2 x = *reinterpret_cast<const double2*>(a.XT + row + 2 * q);
            double2 ax = make_double2(0.0, 0.0);
            if (a.dense_target) ax = *reinterpret_cast<const double2*>(a.AX + row + 2 * q);
            double g0, g1, l0, l1;
            hb_grad(a, 2 * q, x.x, ax.x, g0, l0);
            hb_grad(a, 2 * q + 1, x.y, ax.y, g1, l1);
            lp += l0 + l1;
            if (kick) {
                p.x = fma(c, g0, p.x);
                p.y = fma(c, g1, p.y);
                *reinterpret_cast<double2*>(a.P + row + 2 * q) = p;
            }
        }
        if (PH == 2) {
            k0 = hb_block_sum(k0, red);
            if (tid == 0) a.H0[chain] = a.U0[chain] + 0.5 * k0;
        } else if (last_kick) {
            lp = hb_block_sum(lp, red);
            if (tid == 0) a.U1[chain] = -lp;
        }
    } else if (PH == 3) {
        for (int q = tid; q < npair; q += HB_THREADS) {
            double2 x = *reinterpret_cast<const double2*>(a.XT + row + 2 * q);
            const double2 mp = *reinterpret_cast<const double2*>((a.dense_mass ? a.MP : a.P) + row + 2 * q);
            x.x = fma(a.eps, mp.x, x.x);
            x.y = fma(a.eps, mp.y, x.y);
            *reinterpret_cast<double2*>(a.XT + row + 2 * q) = x;
        }
    } else {   // PH == 5
        double k1 = 0.0;
        for (int q = tid; q < npair; q += HB_THREADS) {
            const double2 p = *reinterpret_cast<const double2*>(a.P + row + 2 * q);
            const double2 mp = a.dense_mass ? *reinterpret_cast<const double2*>(a.MP + row + 2 * q) : p;
            k1 = fma(p.x, mp.x, fma(p.y, mp.y, k1));
        }
        k1 = 0.5 * hb_block_sum(k1, red);
        double U1 = (a.n_leap > 0) ? a.U1[chain] : a.U0[chain];
        if (!isfinite(U1)) U1 = CUDART_INF;                                       // src/hmc.cpp:180-182
        const double comp = fmin(0.01, -(U1 + k1) + a.H0[chain]);                 // :187
        const bool acc = a.Uu[chain] < exp(comp);                                 // :188-191
        const bool keep = t >= a.n_burnin;
        double* out = keep ? a.draws + ((size_t)chain * a.n_keep + (size_t)(t - a.n_burnin)) * d : nullptr;
        for (int q = tid; q < npair; q += HB_THREADS) {
            double2 x;
            if (acc) {
                x = *reinterpret_cast<const double2*>(a.XT + row + 2 * q);
                *reinterpret_cast<double2*>(a.X + row + 2 * q) = x;
            } else {
                x = *reinterpret_cast<const double2*>(a.X + row + 2 * q);
            }
            if (keep) *reinterpret_cast<double2*>(out + 2 * q) = x;
        }
        if (tid == 0) {
            if (acc) a.U0[chain] = U1;
            if (keep) {
                if (acc) a.n_accept[chain] += 1;
                if (a.logp) a.logp[(size_t)chain * a.n_keep + (t - a.n_burnin)] = -(acc ? U1 : a.U0[chain]);
            }
        }
        // the draw index advances once per draw: by the last block to get here it would race with readers, so a separate
        // single-thread kernel does it (hb_next_draw)
    }
}

__global__ void hb_next_draw(int* t_dev) { *t_dev += 1; }
__global__ void hb_set_draw(int* t_dev, int v) { *t_dev = v; }

bool hmc_batched_supported(int target_id, int d, bool has_precond, bool strict, bool has_bounds, long long n_chains)
{
    const bool dense_target = target_id == MCMCB200_TARGET_DENSE_GAUSS || target_id == MCMCB200_TARGET_LINREG;
    const bool elementwise = target_id == MCMCB200_TARGET_ISO_GAUSS || target_id == MCMCB200_TARGET_DIAG_GAUSS;
    if (strict || has_bounds || (d & 1) || d < 2 || d > 2048) return false;
    if (!(dense_target || (elementwise && has_precond))) return false;   // something must be dense for a GEMM to pay
    if (const char* e = std::getenv("MCMCB200_HMC_BATCHED")) return e[0] != '0';
    return d > 32 * MAX_EPL || n_chains >= 512;   // beyond the register-resident kernels, or enough chains to fill GEMM tiles
}

long long hmc_batched_work_doubles(long long n_chains, int d) { return 6 * n_chains * (long long)d + 4 * n_chains + 8; }

int launch_hmc_batched(const HmcLaunch& h, double* work, int* launches)
{
    if (h.broadcast_x0) { set_error("hmc (chain-batched path): broadcast_initial is not supported"); return MCMCB200_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(h.x0) | reinterpret_cast<uintptr_t>(h.draws)) & 15) {
        set_error("hmc (chain-batched path): initial_vals and draws_out must be 16-byte aligned");
        return MCMCB200_ERR_UNSUPPORTED;
    }
    const long long C = h.n_chains;
    const int d = h.d;
    const size_t cd = (size_t)C * d;
    HbArgs a;
    a.n_chains = C; a.d = d; a.target_id = h.target_id; a.tdata = h.tdata; a.chain_offset = h.chain_offset; a.rng = h.rng;
    a.eps = h.eps; a.n_leap = h.n_leap;
    a.dense_mass = h.S_cm != nullptr;
    a.dense_target = (h.target_id == MCMCB200_TARGET_DENSE_GAUSS || h.target_id == MCMCB200_TARGET_LINREG);
    a.x0 = h.x0;
    a.X = work; a.XT = work + cd; a.P = work + 2 * cd;
    a.Z = a.dense_mass ? work + 3 * cd : a.P;
    a.MP = a.dense_mass ? work + 4 * cd : a.P;
    a.AX = work + 5 * cd;
    a.U0 = work + 6 * cd; a.H0 = a.U0 + C; a.U1 = a.H0 + C; a.Uu = a.U1 + C;
    a.t_dev = reinterpret_cast<int*>(a.Uu + C);
    a.n_accept = h.n_accept; a.draws = h.draws; a.logp = h.logp; a.n_keep = h.n_keep; a.n_burnin = h.n_burnin;
    cudaStream_t st = h.stream;
    const bool philox = h.rng.mode == RNG_PHILOX;
    int nl = 0, rc = MCMCB200_OK;
    auto gemm = [&](const double* Y, const double* B, double* Cc) { ++nl; return launch_dgemm_dmma(Y, B, Cc, C, d, st); };
#define HB_ROWS(PH, LAST)                                                                                              \
    do {                                                                                                               \
        if (philox) hmc_batched_rows<PH, RNG_PHILOX><<<(unsigned)C, HB_THREADS, 0, st>>>(a, LAST);                      \
        else hmc_batched_rows<PH, RNG_TAPE><<<(unsigned)C, HB_THREADS, 0, st>>>(a, LAST);                               \
        ++nl;                                                                                                          \
    } while (0)
    // one draw's launch sequence
    auto one_draw = [&]() -> int {
        HB_ROWS(1, 0);
        if (a.dense_mass) {
            if ((rc = gemm(a.Z, h.S_cm, a.P))) return rc;          // p = sqrtM z      (row-major Z times the column-major image of sqrtM)
            if ((rc = gemm(a.P, h.Minv_cm, a.MP))) return rc;      // M^-1 p for K0
        }
        if (a.n_leap > 0) {
            if (a.dense_target && (rc = gemm(a.XT, h.tdata, a.AX))) return rc;
            HB_ROWS(2, 0);
            for (int s = 0; s < a.n_leap; ++s) {
                if (a.dense_mass && (rc = gemm(a.P, h.Minv_cm, a.MP))) return rc;
                HB_ROWS(3, 0);
                if (a.dense_target && (rc = gemm(a.XT, h.tdata, a.AX))) return rc;
                HB_ROWS(4, s + 1 == a.n_leap ? 1 : 0);
            }
            if (a.dense_mass && (rc = gemm(a.P, h.Minv_cm, a.MP))) return rc;
        } else {
            HB_ROWS(2, 0);   // K0 only
        }
        HB_ROWS(5, 0);
        hb_next_draw<<<1, 1, 0, st>>>(a.t_dev);
        ++nl;
        return MCMCB200_OK;
    };
    hb_set_draw<<<1, 1, 0, st>>>(a.t_dev, 0);
    if (a.dense_target && (rc = gemm(h.x0, h.tdata, a.AX))) return rc;
    HB_ROWS(0, 0);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    const long long n_total = h.n_burnin + h.n_keep;
    // capture one draw into a CUDA graph and replay it (the per-draw state — the draw index — is in device memory).
    // Capture needs a non-default stream: with the legacy default stream the launches are issued directly.
    bool graphed = false;
    const int nl_before = nl;
    if (st != nullptr && n_total > 1 && !std::getenv("MCMCB200_NO_GRAPH")) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = one_draw();
            const cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc == MCMCB200_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                const int per_draw = nl - nl_before;
                for (long long t = 0; t < n_total; ++t) MCMCB200_CUDA_TRY(cudaGraphLaunch(exec, st));
                nl = nl_before + (int)(per_draw * n_total);
                graphed = true;
            }
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
            if (!graphed) { cudaGetLastError(); nl = nl_before; if (rc) return rc; }
        } else {
            cudaGetLastError();
        }
    }
    if (!graphed)
        for (long long t = 0; t < n_total; ++t)
            if ((rc = one_draw())) return rc;
#undef HB_ROWS
    MCMCB200_CUDA_TRY(cudaGetLastError());
    *launches = nl;
    return MCMCB200_OK;
}

}  // namespace mcmcb200
