// Chain-batched NUTS: all chains advance in lock-step ROUNDS; a round = one fp64 tensor-core GEMM that evaluates the pending
// gradient product of EVERY chain, followed by one launch of a resumable per-chain state machine (one warp per chain).
//
// Replaces, for dense quadratic targets with many chains (BASELINE config 4: d = 256 dense Gaussian, 4096 chains), the
// persistent kernels of nuts.cu.  Same algorithm — internal::nuts_impl (/root/reference/src/nuts.cpp:30-332),
// nuts_find_initial_step_size (include/mcmc/nuts.ipp:30-93) and the recursive nuts_build_tree (nuts.ipp:97-241) restated
// per SURVEY Appendix C as a memoised walk over the distinct subtrees of a doubling, bug-compatibly (Q12-Q15) — but a
// different execution model:
//
//   * nuts.cu keeps a chain's whole control flow in one warp of a persistent kernel: 255 registers, 8 warps per SM, and the
//     serial tree logic between two gradient products (memo look-ups, U-turn tests on stored states) is a latency chain that
//     nothing hides — ncu: issue slots 17 % busy, the fp64 tensor pipe 13 %, C4 at 0.14 of the fp64 peak.
//   * here a chain is a coroutine.  Its state lives in HBM / L2 (control block, trajectory states, summary table); a step
//     launch resumes every chain at the point where it asked for a product, consumes the GEMM's row, runs the tree logic up
//     to the NEXT product request (the next leapfrog of the lazily extended trajectory), stores its state and returns.
//     Registers are free between rounds (128 per thread, 16 warps per SM instead of 8), and the product of a round is a
//     plain [chains x d] x [d x d] GEMM on the fp64 tensor cores (nuts_ls_gemm below: mma.sync.m8n8k4.f64 — tcgen05 has no
//     f64 kind, DMMA is the fp64 tensor path of sm_100a — 32 x 64 CTA tiles, bit-identical to dgemm_dmma_kernel).
//   * chains do NOT wait for each other across draws: a round advances every unfinished chain by one product, whatever
//     draw / doubling it is in, so the number of rounds is the largest per-chain product count, not the sum of per-draw
//     maxima (C4: 63 k rounds for 4096 chains x 400 draws x 143 leapfrogs).
//   * the chains of a call are split into GROUPS of ~512, each on its own stream with its rounds captured in a CUDA graph:
//     a group's rounds are a dependent sequence of short kernels (the round's latency, ~25 us, is what a lone group runs at),
//     different groups overlap — one group's GEMM on the tensor pipe under the other groups' step kernels.
//   * what a resume costs was cut down step by step (ncu + clock64 instrumentation, MCMCB200_DEBUG=1): the control block and
//     the summary table arrive by cp.async straight into shared memory together with the three vectors of the pending
//     leapfrog (one exposed latency); the U-turn tests of every subtree that ends in a state are evaluated when the state
//     is created (the DAG's (level, state) pairs are tabulated on the host) and recorded as bits, so the walk is scalar and
//     touches shared memory only; the next leapfrog's first half is staged speculatively right away; a finished draw's
//     bookkeeping + momentum refresh is handed to the next round instead of stretching the launch for all other chains.
//
// Products requested per chain: one per distinct trajectory state (as nuts.cu), one for the gradient at prev_draw whenever
// prev_draw changed (nuts.cu re-evaluates it at every doubling; the value is cached here), the initial-step-size search.
// Scope: FAST arithmetic (the GEMM accumulates in tensor-core order: held to the contract tolerance, not to bits),
// targets dense_gauss / linreg, M = I, no box constraints, n_dim even and <= 512, Philox or a caller tape; everything else
// stays on nuts.cu.  MCMCB200_NUTS_BATCHED=0/1 forces the choice.
#include "engine.h"
#include "rng.cuh"
#include "dgemm.h"
#include "nuts_batched.h"
#include <math_constants.h>
#include <cstdlib>
#include <cstdio>
#include <vector>

namespace mcmcb200
{

namespace
{

constexpr int LS_WARPS = 4;            // chains per CTA of the step kernel
constexpr int LS_LEVELS = 22;          // max_tree_depth <= 20
constexpr int LS_GRAPH_ROUNDS = 32;    // rounds captured in one CUDA graph (per group)
constexpr int LS_POLL_GRAPHS = 8;      // graph launches per group between two reads of the running-chain counters
constexpr int LS_MAX_GROUPS = 16;
constexpr unsigned LS_EPOCH_MAX = 2047u;
constexpr int LS_TAB_SMEM = 256;       // summary-table entries per chain kept in shared memory: covers max_tree_depth <= 10 (249)
constexpr int LS_STATES_SMEM = 64;     // per-state leaf records kept in the control block: covers max_tree_depth <= 11 (56 states)

__host__ __device__ inline int ls_m_max(int max_depth)
{
    const int j = max_depth > 0 ? max_depth - 1 : 0;   // deepest tree built is depth max_depth - 1
    return 1 + j * (j + 1) / 2;
}
__host__ __device__ inline int ls_memo_entries(int max_depth)
{
    const int Jm = max_depth - 1;
    int n = 0;
    for (int j = 1; j <= Jm; ++j) n += (Jm * (Jm + 1) - j * (j + 1)) / 2 + 1;
    return n;
}

// summary of a built subtree (same packing as nuts.cu): w0 = n (21 bits) | far slot << 21 (8 bits) | s << 29,
// w1 = n_alpha (21 bits) | epoch << 21 (11 bits, 0 = empty)
struct LsSummary { double alpha; unsigned w0, w1; };

enum { PH_INIT0 = 0, PH_INIT_GRAD, PH_INIT_LF, PH_DRAW_BEGIN, PH_DBL_BEGIN, PH_DBL_G0, PH_WALK_INIT, PH_LEAF, PH_WALK, PH_DRAW_END, PH_DONE, PH_DBL_END };

// a chain's control block: everything the coroutine needs besides its vectors (which sit in the chain's work area)
struct LsCtl {
    int ph, t, depth, dir, computed, level, ucount, n_alpha, good_round, s_val, n_acc, g0_valid, R_n, R_s, R_nalpha, R_far;
    unsigned epoch;
    int staged;   // the first half of the NEXT leapfrog has been staged speculatively (TX / TRh hold it)
    long long n_val, n_lf, cursor, ubase_cur;
    double eps, mu, h, eps_bar, prev_U, prev_K, log_u, alpha, R_alpha, H0, e_signed, pU, pK, lp0;
    // traversal stack: frame = (j, a, phase) + the first half's summary while the second is built
    int sj[LS_LEVELS], sa[LS_LEVELS], sph[LS_LEVELS], sn[LS_LEVELS], snalpha[LS_LEVELS];
    double salpha[LS_LEVELS];
    // per trajectory state k (slot k - 1) of the running doubling, written when the state is created:
    //   ut: bit j (1 <= j <= depth) = outcome of the U-turn test between states k - j and k (the ends of T(j, k - j - 1));
    //       bit 30 = n, bit 31 = s of the leaf (nuts.ipp:146-147);   lalpha = the leaf's alpha statistic (:157)
    unsigned ut[LS_STATES_SMEM];
    double lalpha[LS_STATES_SMEM];
};
static_assert(sizeof(LsCtl) % 16 == 0, "control block is copied as 128-bit words");
constexpr int LS_CTL_WORDS = (int)(sizeof(LsCtl) / 8);

struct LsArgs {
    long long n_chains;
    int d;
    int target_id;
    const double* tdata;
    const double* x0;
    int broadcast_x0;
    long long chain_offset;
    RngArgs rng;
    double* draws;
    double* logp;
    long long* n_accept;
    double* step_out;
    long long* n_leapfrog;
    long long* tape_used;
    long long n_burnin, n_keep, n_adapt, t_end;
    int max_depth;
    double eps_bar0, delta, gamma, t0, kappa;
    const double2* tab;      // Box-Muller tables (rng.cuh), built once per run in global memory
    double* TX;              // [C][d] the position each chain wants multiplied (GEMM input)
    double* TY;              // [C][d] the products (GEMM output)
    LsCtl* ctl;              // [C]
    double* work;            // per-chain areas
    long long work_stride;
    int* n_running;          // this group's running-chain counter
    int* next_chain;         // persistent kernel: the next chain to hand to a free warp
    int split_dbl_end;       // hand the closing work of a doubling to the next round (MCMCB200_NUTS_SPLIT=0 keeps it in the leaf's resume)
    long long chain_base;    // this launch covers chains [chain_base, chain_base + n_group) of the call
    long long n_group;
    const unsigned* jmask;   // [max_depth][m_max + 1]: for a doubling of depth D and state k, the levels j whose subtree T(j, k - j - 1) exists
    int jm_stride;
    int m_max, memo_n;       // ls_m_max / ls_memo_entries of max_depth
    int lvl_off[LS_LEVELS];  // off(j) of the summary table: level j starts after the levels above it
    unsigned long long* dbg; // MCMCB200_DEBUG: per path category {cycles, count, max cycles} of a warp's resume (else null)
    unsigned* ut_g;          // [C][m_max] / [C][m_max]: the per-state records when they do not fit the control block
    double* lalpha_g;
};

// per-chain work area (doubles): prev_draw, draw momentum, theta+, theta-, r+, r-, half-kicked momentum of the pending
// leapfrog, gradient at prev_draw (8 d), states k = 1..m_max as (x_k, r_k), U_k[m_max], K_k[m_max], summary table
__host__ __device__ inline long long ls_work_per_chain(int d, int max_depth)
{
    const long long m = ls_m_max(max_depth), e = ls_memo_entries(max_depth);
    long long n = 8ll * d + 2ll * d * m + 2 * m + 2 * e;
    return (n + 1) & ~1ll;
}

template <int EPL, bool FT = false> __device__ __forceinline__ void ldv(const double* src, int d, int lane, double (&x)[EPL])
{
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m) {
        const int j = m * 64 + 2 * lane;
        if (FT || j < d) {
            const double2 v = __ldcg(reinterpret_cast<const double2*>(src + j));
            x[2 * m] = v.x;
            x[2 * m + 1] = v.y;
        } else {
            x[2 * m] = 0.0;
            x[2 * m + 1] = 0.0;
        }
    }
}
__device__ __forceinline__ long long ls_clock()   // ordered against barriers and memory operations (the intrinsic is not)
{
    long long v;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(v)::"memory");
    return v;
}
__device__ __forceinline__ void ls_cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
template <int EPL, bool FT = false> __device__ __forceinline__ void stv(double* dst, int d, int lane, const double (&x)[EPL])
{
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m) {
        const int j = m * 64 + 2 * lane;
        if (FT || j < d) *reinterpret_cast<double2*>(dst + j) = make_double2(x[2 * m], x[2 * m + 1]);
    }
}

// log pi(x) and its gradient from x and the product y = A x (targets.cuh: DenseGauss, LinReg; FAST arithmetic).
// Returns this lane's partial sum; the warp total goes through ls_value.
template <int EPL, bool WANT_VALUE>
__device__ __forceinline__ double ls_eval_lane(const LsArgs& a, int lane, const double (&x)[EPL], const double (&y)[EPL], double (&g)[EPL])
{
    if (a.target_id == MCMCB200_TARGET_DENSE_GAUSS) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) g[k] = -y[k];
        return WANT_VALUE ? lane_dot<EPL, false>(x, y) : 0.0;
    }
    const double* __restrict__ bp = a.tdata + (size_t)a.d * (size_t)a.d;   // linreg: log pi = x.(b - A x / 2), grad = b - A x
    double t[EPL];
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        const int j = elem_index(lane, k);
        const double b = (j < a.d) ? __ldg(bp + j) : 0.0;
        g[k] = b - y[k];
        t[k] = b - 0.5 * y[k];
    }
    return WANT_VALUE ? lane_dot<EPL, false>(x, t) : 0.0;
}
__device__ __forceinline__ double ls_value(const LsArgs& a, double warp_total)
{
    return (a.target_id == MCMCB200_TARGET_DENSE_GAUSS) ? -(0.5 * warp_total) : warp_total;
}
template <int EPL, bool WANT_VALUE>
__device__ __forceinline__ double ls_eval(const LsArgs& a, int lane, const double (&x)[EPL], const double (&y)[EPL], double (&g)[EPL])
{
    const double sl = ls_eval_lane<EPL, WANT_VALUE>(a, lane, x, y, g);
    return WANT_VALUE ? ls_value(a, warp_sum<false>(sl)) : 0.0;
}

__device__ __forceinline__ double ls_neg_logp_finite(double lp)
{
    const double U = -lp;
    return isfinite(U) ? U : CUDART_INF;   // "if (!std::isfinite(prop_U)) prop_U = posinf"
}

// ------------------------------------------------------------------------------------------------ the round's GEMM
// TY[M x d] = TX[M x d] * A[d x d] (row-major, A symmetric) with mma.sync.m8n8k4.f64.  Same fragment layout and the same
// accumulation order over k as dgemm_dmma_kernel (mala_wide.cu) — the results are bit-identical to it — but a 32 x 64 CTA tile
// of 4 warps instead of 128 x 64 of 8: a round's GEMM is short (a group's 1024 chains x 256 x 256), so what matters is its
// LATENCY — many small CTAs spread over all SMs, several resident per SM next to the step kernels of the other groups.
constexpr int SG_M = 32, SG_N = 64, SG_K = 16, SG_THREADS = 128, SG_STAGES = 3;
constexpr int SG_ASTR = SG_K + 4, SG_BSTR = SG_N + 4;   // padded strides: conflict-free 64-bit fragment loads
constexpr int SG_STAGE_DOUBLES = SG_M * SG_ASTR + SG_K * SG_BSTR;

__device__ __forceinline__ void sg_dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(SG_THREADS, 4) nuts_ls_gemm(const double* __restrict__ Y, const double* __restrict__ Amat, double* __restrict__ Cout, int M, int d)
{
    __shared__ __align__(16) double gsm[SG_STAGES * SG_STAGE_DOUBLES];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * SG_M, n0 = blockIdx.x * SG_N;
    const int nk = (d + SG_K - 1) / SG_K;
    auto load_stage = [&](int stage, int kt) {
        double* As = gsm + stage * SG_STAGE_DOUBLES;
        double* Bs = As + SG_M * SG_ASTR;
        const int k0 = kt * SG_K;
#pragma unroll
        for (int cidx = tid; cidx < SG_M * (SG_K / 2); cidx += SG_THREADS) {
            const int r = cidx / (SG_K / 2), cc = (cidx % (SG_K / 2)) * 2;
            double* dst = As + r * SG_ASTR + cc;
            const int gr = m0 + r, gc = k0 + cc;
            if (gr < M && gc + 1 < d) ls_cp_async16(dst, Y + (size_t)gr * d + gc);
            else { dst[0] = (gr < M && gc < d) ? Y[(size_t)gr * d + gc] : 0.0; dst[1] = 0.0; }
        }
#pragma unroll
        for (int cidx = tid; cidx < SG_K * (SG_N / 2); cidx += SG_THREADS) {
            const int r = cidx / (SG_N / 2), cc = (cidx % (SG_N / 2)) * 2;
            double* dst = Bs + r * SG_BSTR + cc;
            const int gr = k0 + r, gc = n0 + cc;
            if (gr < d && gc + 1 < d) ls_cp_async16(dst, Amat + (size_t)gr * d + gc);
            else { dst[0] = (gr < d && gc < d) ? Amat[(size_t)gr * d + gc] : 0.0; dst[1] = 0.0; }
        }
    };
    double acc[4][2][2];   // warp tile: all 32 rows x 16 columns (columns n0 + 16 warp ...)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int s = 0; s < SG_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        asm volatile("cp.async.commit_group;\n" ::);
    }
    for (int kt = 0; kt < nk; ++kt) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(SG_STAGES - 2));
        __syncthreads();
        {
            const int nxt = kt + SG_STAGES - 1;
            if (nxt < nk) load_stage(nxt % SG_STAGES, nxt);
            asm volatile("cp.async.commit_group;\n" ::);
        }
        const double* As = gsm + (kt % SG_STAGES) * SG_STAGE_DOUBLES;
        const double* Bs = As + SG_M * SG_ASTR;
#pragma unroll
        for (int kk = 0; kk < SG_K / 4; ++kk) {
            double af[4], bf[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[(i * 8 + g) * SG_ASTR + kk * 4 + t];
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[j] = Bs[(kk * 4 + t) * SG_BSTR + warp * 16 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) sg_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = n0 + warp * 16 + j * 8 + 2 * t;
            if (row < M && col + 1 < d) *reinterpret_cast<double2*>(Cout + (size_t)row * d + col) = make_double2(acc[i][j][0], acc[i][j][1]);
            else if (row < M && col < d) Cout[(size_t)row * d + col] = acc[i][j][0];
        }
    }
}

__global__ void nuts_ls_tables(double2* tab) { build_rng_tables(tab); }

__global__ void nuts_ls_init(LsCtl* ctl, long long n_chains, int* n_running, int n_groups, long long group_chains)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_groups) {
        const long long left = n_chains - c * group_chains;
        n_running[c] = (int)(left < group_chains ? left : group_chains);
    }
    if (c >= n_chains) return;
    unsigned long long* w = reinterpret_cast<unsigned long long*>(ctl + c);
    for (int i = 0; i < LS_CTL_WORDS; ++i) w[i] = 0ull;   // ph = PH_INIT0
}

// A chain's resume: from the arrival of the product it asked for to its next product request (or the end of its run).
//   c / memo : control block and (MEMO_SH) summary table in shared memory;  memo_g: the table's home in global memory
//   TX / TRh : where the next request's position / half-kicked momentum are staged (global rows, or shared memory in the
//              persistent kernel);  xt, rt, yin: the pending position, half-kicked momentum and the product, already loaded
//   PERSIST  : the control block and table stay on chip for the chain's whole run (no write-through, see nuts_pc_kernel)
template <int EPL, int RNGM, bool MEMO_SH, bool FT, bool PERSIST>
__device__ __forceinline__ void ls_resume(const LsArgs& a, const long long chain, const int lane, LsCtl& c, LsSummary* const memo,
                                          LsSummary* const memo_g, double* const TX, double* const TRh, double (&xt)[EPL], double (&rt)[EPL],
                                          const double (&yin)[EPL])
{
    const int* const lvl_off = a.lvl_off;   // kernel-parameter (constant) bank, indexed dynamically
    const int d = a.d;
    const int m_max = a.m_max;
    const int memo_n = a.memo_n;
    double* const W = a.work + chain * a.work_stride;
    double* const Wprev = W;
    double* const Wm = W + d;
    double* const Wxp = W + 2 * d;
    double* const Wxn = W + 3 * d;
    double* const Wrp = W + 4 * d;
    double* const Wrn = W + 5 * d;
    double* const G0 = W + 7 * d;
    double* const Wst = W + 8 * d;   // state k (1-based): x at Wst + (k-1)*2d, r at + d
    double* const Us = Wst + (size_t)2 * d * m_max;
    double* const Ks = Us + m_max;
    double gt[EPL];
    const long long dbg_t0 = a.dbg ? clock64() : 0;
    long long dbg_t1 = 0, dbg_t2 = 0, dbg_t3 = 0;
    const int dbg_ph = c.ph, dbg_t = c.t, dbg_depth = c.depth;   // (dbg_ph also steers the draw-boundary split)
    // per-state leaf records: in the control block when they fit, else in global memory
    unsigned* const utp = MEMO_SH ? c.ut : a.ut_g + chain * m_max;   // (MEMO_SH: max_tree_depth <= 10, at most 46 states)
    double* const lalp = MEMO_SH ? c.lalpha : a.lalpha_g + chain * m_max;

    // scalars: the control block in shared memory is "warp-uniform memory" — every lane reads and writes the same words with the
    // same values (a same-value store from all lanes is one transaction), so each lane only ever depends on its own accesses
    // and no __syncwarp is needed around them; the compiler keeps what is hot in registers
    int& staged = c.staged;
    int &ph = c.ph, &t = c.t, &depth = c.depth, &dir = c.dir, &computed = c.computed, &level = c.level, &ucount = c.ucount, &n_alpha = c.n_alpha,
        &good_round = c.good_round, &s_val = c.s_val, &n_acc = c.n_acc, &g0_valid = c.g0_valid, &R_n = c.R_n, &R_s = c.R_s, &R_nalpha = c.R_nalpha,
        &R_far = c.R_far;
    unsigned& epoch = c.epoch;
    long long &n_val = c.n_val, &n_lf = c.n_lf, &ubase_cur = c.ubase_cur;
    double &eps = c.eps, &mu = c.mu, &h = c.h, &eps_bar = c.eps_bar, &prev_U = c.prev_U, &prev_K = c.prev_K, &log_u = c.log_u, &alpha = c.alpha,
           &R_alpha = c.R_alpha, &H0 = c.H0, &e_signed = c.e_signed, &pU = c.pU, &pK = c.pK, &lp0 = c.lp0;
    ChainRng<RNGM> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);
    rng.cursor = c.cursor;

    auto kinetic = [&](const double (&p)[EPL]) -> double { return 0.5 * warp_dot<EPL, false>(p, p); };   // K = p.p/2 (src/nuts.cpp:204, nuts.ipp:140)
    auto kick = [&](double e, double (&p)[EPL], const double (&g)[EPL]) {   // half kick p + (e grad)/2 (src/nuts.cpp:139-154)
#pragma unroll
        for (int k = 0; k < EPL; ++k) p[k] = fma(0.5 * e, g[k], p[k]);
    };
    // first half of a leapfrog of (signed) size e from the tip (xt, rt, gt): half kick, drift, post the product request
    auto begin_leapfrog = [&](double e, double (&xt)[EPL], double (&rt)[EPL], const double (&gt)[EPL]) {
        kick(e, rt, gt);
#pragma unroll
        for (int k = 0; k < EPL; ++k) xt[k] = fma(e, rt[k], xt[k]);
        stv<EPL>(TX, d, lane, xt);
        stv<EPL>(TRh, d, lane, rt);
    };
    // leaf T(0, ao): state k = ao + 1 of the trajectory (nuts.ipp:132-157)
    struct Leaf { int n, s; };
    auto leaf_ns = [&](int k_state) -> Leaf {
        const unsigned w = utp[k_state - 1];
        Leaf l;
        l.n = (int)((w >> 30) & 1u);   // :146, evaluated when the state was created
        l.s = (int)(w >> 31);          // :147
        return l;
    };

    bool yielded = false;
    while (!yielded) {
        switch (ph) {
        case PH_INIT0: {
            double x[EPL];
            load_vec<EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), d, lane, x);
            stv<EPL>(Wprev, d, lane, x);
            stv<EPL>(TX, d, lane, x);
            {
                LsSummary* const mz = PERSIST ? memo : memo_g;   // (the launched kernel reloads its working copy from the home every round)
                for (int i = lane; i < memo_n; i += 32) { mz[i].alpha = 0.0; mz[i].w0 = 0u; mz[i].w1 = 0u; }
            }
            ph = PH_INIT_GRAD;
            yielded = true;
            break;
        }
        case PH_INIT_GRAD: {
            // pre-loop momentum draw (src/nuts.cpp:166-168) and nuts_find_initial_step_size (nuts.ipp:30-93)
            lp0 = ls_eval<EPL, true>(a, lane, xt, yin, gt);   // TX holds x0 (PH_INIT0)
            stv<EPL>(G0, d, lane, gt);
            g0_valid = 1;
            pU = ls_neg_logp_finite(lp0);
            rng.template normals<EPL, false>(a.rng, -1, d, lane, a.tab, rt);
            pK = kinetic(rt);
            eps = 1.0;
            begin_leapfrog(eps, xt, rt, gt);
            ph = PH_INIT_LF;
            yielded = true;
            break;
        }
        case PH_INIT_LF: {
            const double qU = ls_neg_logp_finite(ls_eval<EPL, true>(a, lane, xt, yin, gt));
            kick(eps, rt, gt);
            ++n_lf;
            const double qK = kinetic(rt);
            const double dH = -(qU + qK) + (pU + pK);
            const int a_val = 2 * (dH > -0.69314718055994530942) - 1;   // > std::log(0.5)
            const bool cond = dH > -0.69314718055994530942;             // > -std::log(2)
            if (cond) {
                eps *= (a_val == 1) ? 2.0 : 0.5;                        // step_size *= std::pow(2, a_val); state NOT reset (Q14)
                begin_leapfrog(eps, xt, rt, gt);
                yielded = true;
                break;
            }
            mu = log(10.0 * eps);   // src/nuts.cpp:174
            h = 0.0;
            eps_bar = a.eps_bar0;
            prev_U = -lp0;          // :181 (no finite clamp here)
            t = 0;
            ph = PH_DRAW_BEGIN;
            break;
        }
        case PH_DRAW_BEGIN: {
            if (t >= (int)a.t_end) {
                if (lane == 0) {
                    if (a.n_accept) a.n_accept[chain] = n_acc;
                    if (a.step_out) a.step_out[chain] = eps;
                    if (a.n_leapfrog) a.n_leapfrog[chain] = n_lf;
                    if (a.tape_used) a.tape_used[chain] = rng.cursor;
                    atomicSub(a.n_running, 1);
                }
                ph = PH_DONE;
                yielded = true;
                break;
            }
            ucount = 0;
            double x[EPL];
            rng.template normals<EPL, false>(a.rng, t, d, lane, a.tab, rt);   // :200
            ubase_cur = rng.cursor;
            prev_K = kinetic(rt);                                             // :204
            stv<EPL>(Wm, d, lane, rt);
            log_u = (log(rng.uniform_at(a.rng, t, ucount++, ubase_cur)) - prev_U) - prev_K;   // :206
            ldv<EPL>(Wprev, d, lane, x);
            stv<EPL>(Wxp, d, lane, x);   // :212-215
            stv<EPL>(Wxn, d, lane, x);
            stv<EPL>(Wrp, d, lane, rt);
            stv<EPL>(Wrn, d, lane, rt);
            depth = 0; s_val = 1; n_alpha = 0; good_round = 0; n_val = 1; alpha = 0.0;
            ph = PH_DBL_BEGIN;
            break;
        }
        case PH_DBL_BEGIN: {
            if (!(s_val == 1 && depth < a.max_depth)) {   // :227
                // the draw is complete.  Its bookkeeping + the next draw's momentum refresh is a long path; a warp that has
                // just finished a state and walked the tree hands it to the next round (which costs this chain one idle
                // GEMM row per draw) instead of stretching this launch for every other chain of the group
                ph = PH_DRAW_END;
                yielded = (dbg_ph == PH_LEAF || dbg_ph == PH_DBL_END);
                break;
            }
            const double zz = rng.uniform_at(a.rng, t, ucount++, ubase_cur);        // :233
            dir = (zz <= 0.5) ? -1 : 1;
            e_signed = (dir == 1) ? eps : -eps;
            H0 = prev_U + prev_K;
            computed = 0;
            // new restart point / direction: the summaries of the previous doubling are void
            if (++epoch > LS_EPOCH_MAX) {
                for (int i = lane; i < memo_n; i += 32) { if (!PERSIST) memo_g[i].w1 = 0u; if (MEMO_SH) memo[i].w1 = 0u; }   // home and working copy
                __syncwarp();
                epoch = 1;
            }
            if (!g0_valid) {   // prev_draw moved since its gradient was last evaluated: one product
                double x[EPL];
                ldv<EPL>(Wprev, d, lane, x);
                stv<EPL>(TX, d, lane, x);
                ph = PH_DBL_G0;
                yielded = true;
                break;
            }
            ph = PH_WALK_INIT;
            break;
        }
        case PH_DBL_G0: {
            ls_eval<EPL, false>(a, lane, xt, yin, gt);   // TX holds prev_draw (PH_DBL_BEGIN)
            stv<EPL>(G0, d, lane, gt);
            g0_valid = 1;
            ph = PH_WALK_INIT;
            break;
        }
        case PH_WALK_INIT: {
            // tip of the trajectory LF^k (prev_draw, draw momentum): every doubling restarts here (Q12)
            ldv<EPL>(Wprev, d, lane, xt);
            ldv<EPL>(Wm, d, lane, rt);
            ldv<EPL>(G0, d, lane, gt);
            R_n = 0; R_s = 0; R_nalpha = 0; R_far = 0; R_alpha = 0.0;
            level = 0;
            c.sj[0] = depth; c.sa[0] = 0; c.sph[0] = 0;
            // state 1 is always needed (the walk's first leaf): post its product now — the walk itself never touches a vector
            begin_leapfrog(e_signed, xt, rt, gt);
            staged = 0;
            ph = PH_LEAF;
            yielded = true;
            break;
        }
        case PH_LEAF: {
            // second half of the leapfrog whose product has arrived: gradient / log pi from the product, second half kick
            {
                // (register copies: stores into the control block's arrays would otherwise force the scalars to be re-read)
                const double es = e_signed, lu = log_u, h0 = H0;
                const int dep = depth, dr = dir, k = computed + 1;
                double sl = ls_eval_lane<EPL, true>(a, lane, xt, yin, gt);
                kick(es, rt, gt);
                double kl = lane_dot<EPL, false>(rt, rt);
                warp_sum2<false>(sl, kl);
                computed = k; ++n_lf;
                const double Uk = ls_neg_logp_finite(ls_value(a, sl));   // :134-138
                const double Kk = 0.5 * kl;                              // :140
                double* const sk = Wst + (size_t)(k - 1) * 2 * d;
                stv<EPL, FT>(sk, d, lane, xt);
                stv<EPL, FT>(sk + d, d, lane, rt);
                if (lane == 0) { Us[k - 1] = Uk; Ks[k - 1] = Kk; }
                // U-turn tests of every subtree that ends in this state: T(j, ao) has near = ao + 1, far = ao + j + 1 (Appendix C),
                // so the tests with far = k pair state k with the states k - j of the levels j the doubling's DAG contains
                // (jmask, built on the host).  The pairs are independent of the walk, which then only looks at the recorded bits.
                unsigned jm = __ldg(a.jmask + dep * a.jm_stride + k);
                // the trajectory continues unless this doubling stops or is complete: stage the next leapfrog's first half NOW
                // (speculatively — a staged state the walk does not ask for is simply overwritten by the next request), so that
                // the gradient is dead before the U-turn tests and the walk needs no vector at all
                const int stg = (k < 1 + dep * (dep + 1) / 2) ? 1 : 0;
                staged = stg;
                if (stg) {
#pragma unroll
                    for (int m = 0; m < EPL / 2; ++m) {
                        const int j = m * 64 + 2 * lane;
                        if (FT || j < d) {
                            const double rh0 = fma(0.5 * es, gt[2 * m], rt[2 * m]), rh1 = fma(0.5 * es, gt[2 * m + 1], rt[2 * m + 1]);
                            *reinterpret_cast<double2*>(TRh + j) = make_double2(rh0, rh1);
                            *reinterpret_cast<double2*>(TX + j) = make_double2(fma(es, rh0, xt[2 * m]), fma(es, rh1, xt[2 * m + 1]));
                        }
                    }
                }
                // the leaf's record (nuts.ipp:146-157)
                const unsigned ln = (lu <= (-Uk - Kk)) ? 1u : 0u;
                const unsigned ls = (lu < ((1000.0 - Uk) - Kk)) ? 1u : 0u;
                lalp[k - 1] = exp(fmin(0.0, -(Uk + Kk) + h0));
                if (a.dbg) dbg_t1 = clock64();
                unsigned bits = (ln << 30) | (ls << 31);
                // (theta_far - theta_near).r for both ends; with dir = -1 the roles of the ends swap, which negates both dot
                // products exactly (pos - neg = -(far - near)): the sign test is applied to dir * d   (:226-229)
                auto test = [&](int j, double (&xn_)[EPL], const double (&rn_)[EPL]) {
#pragma unroll
                    for (int q = 0; q < EPL; ++q) xn_[q] = xt[q] - xn_[q];
                    double d_near = lane_dot<EPL, false>(xn_, rn_);
                    double d_far = lane_dot<EPL, false>(xn_, rt);
                    warp_sum2<false>(d_near, d_far);
                    const bool ok = (dr == 1) ? (d_near >= 0.0 && d_far >= 0.0) : (d_near <= 0.0 && d_far <= 0.0);
                    if (ok) bits |= 1u << j;
                };
                while (jm) {
                    const int j = __ffs(jm) - 1;
                    jm &= jm - 1;
                    double xn_[EPL], rn_[EPL];
                    ldv<EPL, FT>(sk - (size_t)j * 2 * d, d, lane, xn_);
                    ldv<EPL, FT>(sk - (size_t)j * 2 * d + d, d, lane, rn_);
                    test(j, xn_, rn_);
                }
                utp[k - 1] = bits;
                if (a.dbg) dbg_t2 = clock64();
            }
            ph = PH_WALK;
        }
        // fall through
        case PH_WALK: {
            // ---- summary of T(depth, 0): explicit-stack walk over the DAG of distinct subtrees (nuts.cu) ----
            {
                int lv = level, rn = R_n, rs = R_s, rna = R_nalpha, rfar = R_far;
                double ralpha = R_alpha;
                const int comp = computed;
                const unsigned ep = epoch;
                while (lv >= 0) {
                    const int j = c.sj[lv], ao = c.sa[lv], sp = c.sph[lv];
                    if (j == 0) {
                        const int k_need = ao + 1;
                        if (comp < k_need) {   // extend the trajectory by one leapfrog (nuts.ipp:132): its product was requested
                            ph = PH_LEAF;       // when the previous state was created (staged is set whenever a further state can be needed)
                            yielded = true;
                            break;
                        }
                        const unsigned w = utp[k_need - 1];
                        rn = (int)((w >> 30) & 1u);   // :146
                        rs = (int)(w >> 31);          // :147
                        ralpha = lalp[k_need - 1];    // :157
                        rna = 1;
                        rfar = k_need;
                        --lv;
                        continue;
                    }
                    LsSummary* const ent = memo + lvl_off[j] + ao;
                    if (sp == 0) {
                        const unsigned w1 = MEMO_SH ? ent->w1 : reinterpret_cast<volatile unsigned*>(&ent->w1)[0];
                        if ((w1 >> 21) == ep) {   // built earlier in this doubling
                            const unsigned w0 = MEMO_SH ? ent->w0 : reinterpret_cast<volatile unsigned*>(&ent->w0)[0];
                            ralpha = MEMO_SH ? ent->alpha : reinterpret_cast<volatile double*>(&ent->alpha)[0];
                            rn = (int)(w0 & 0x1fffffu); rfar = (int)((w0 >> 21) & 0xffu); rs = (int)((w0 >> 29) & 1u);
                            rna = (int)(w1 & 0x1fffffu);
                            --lv;
                        } else {
                            c.sph[lv] = 1; c.sj[lv + 1] = j - 1; c.sa[lv + 1] = ao; c.sph[lv + 1] = 0;
                            ++lv;
                        }
                        continue;
                    }
                    if (sp == 1 && rs == 1) {   // first half returned and did not stop: build the second half from far(A)
                        c.sn[lv] = rn; c.salpha[lv] = ralpha; c.snalpha[lv] = rna;
                        c.sph[lv] = 2;
                        c.sj[lv + 1] = j - 1; c.sa[lv + 1] = ao + j; c.sph[lv + 1] = 0;
                        ++lv;
                        continue;
                    }
                    if (sp == 2) {   // second half returned
                        // U-turn test on the merged subtree's ends: near = ao+1, far = ao+j+1 (Appendix C), evaluated when state
                        // `far` was created; its outcome only matters while the subtree has not stopped (s = s'' * ..., nuts.ipp:229)
                        const int far = ao + j + 1;
                        if (rs == 1) rs = (int)((utp[far - 1] >> j) & 1u);
                        rn = c.sn[lv] + rn;
                        ralpha = c.salpha[lv] + ralpha;
                        rna = c.snalpha[lv] + rna;
                        rfar = far;
                    }
                    // (sp == 1 with rs == 0: the result is the first half's, nuts.ipp:234-239)
                    const unsigned nw0 = (unsigned)rn | ((unsigned)rfar << 21) | ((unsigned)rs << 29), nw1 = (unsigned)rna | (ep << 21);
                    if (MEMO_SH) {   // working copy in shared memory (a same-value store from every lane) + write-through to its home
                        ent->alpha = ralpha; ent->w0 = nw0; ent->w1 = nw1;
                        if (!PERSIST && lane == 0) *reinterpret_cast<double2*>(memo_g + (ent - memo)) = make_double2(ralpha, __hiloint2double((int)nw1, (int)nw0));
                    } else {
                        if (lane == 0) { ent->alpha = ralpha; ent->w0 = nw0; ent->w1 = nw1; }
                        __syncwarp();
                    }
                    --lv;
                }
                level = lv; R_n = rn; R_s = rs; R_nalpha = rna; R_far = rfar; R_alpha = ralpha;
            }
            if (a.dbg) dbg_t3 = clock64();
            if (yielded) break;
            // The doubling's tree is complete.  Its closing work (far slot, theta' selection with one Philox block per level,
            // the trajectory's U-turn test: four more vector loads) would make this resume the longest of its launch / CTA
            // round; MCMCB200_NUTS_SPLIT=1 hands it to the next round (one idle product row per doubling, ~5 % more rounds).
            // Measured on C4: 3 % slower in both the launched and the persistent variant, so it is off by default.
            ph = PH_DBL_END;
            if (a.split_dbl_end) { yielded = true; break; }
        }
        // fall through
        case PH_DBL_END: {
            alpha = R_alpha;   // overwritten by every doubling (Q12)
            n_alpha = R_nalpha;
            const int ubase = ucount;   // the merges of this doubling drew uniforms ubase .. ubase + n_alpha - 2 (post-order)
            ucount += R_nalpha - 1;
            // the far slot of T lands in theta^v / r^v (src/nuts.cpp:241-256)
            double xf_[EPL], rf_[EPL];
            ldv<EPL>(Wst + (size_t)(R_far - 1) * 2 * d, d, lane, xf_);
            ldv<EPL>(Wst + (size_t)(R_far - 1) * 2 * d + d, d, lane, rf_);
            stv<EPL>(dir == 1 ? Wxp : Wxn, d, lane, xf_);
            stv<EPL>(dir == 1 ? Wrp : Wrn, d, lane, rf_);
            if (R_s == 1) {
                const double z3 = rng.uniform_at(a.rng, t, ucount++, ubase_cur);   // :261
                if (z3 < (double)R_n / (double)n_val) {                            // :263
                    // theta' of T(depth, 0), resolved lazily: at every merge on the way down the second half's theta' replaces
                    // the first half's with probability n''/(n' + n'') (nuts.ipp:213-221), the merge's uniform being the one
                    // drawn after both halves were built
                    int jj = depth, aa = 0, ob = ubase;
                    while (jj > 0) {
                        int nA, sA, cA, nB, cB;
                        if (jj == 1) {
                            const Leaf lA = leaf_ns(aa + 1);
                            nA = lA.n; sA = lA.s; cA = 0;
                        } else {
                            const LsSummary* eA = memo + lvl_off[jj - 1] + aa;
                            const unsigned w0 = reinterpret_cast<const volatile unsigned*>(&eA->w0)[0], w1 = reinterpret_cast<const volatile unsigned*>(&eA->w1)[0];
                            nA = (int)(w0 & 0x1fffffu); sA = (int)((w0 >> 29) & 1u); cA = (int)(w1 & 0x1fffffu) - 1;
                        }
                        if (sA == 1) {
                            if (jj == 1) {
                                nB = leaf_ns(aa + jj + 1).n; cB = 0;
                            } else {
                                const LsSummary* eB = memo + lvl_off[jj - 1] + aa + jj;
                                nB = (int)(reinterpret_cast<const volatile unsigned*>(&eB->w0)[0] & 0x1fffffu);
                                cB = (int)(reinterpret_cast<const volatile unsigned*>(&eB->w1)[0] & 0x1fffffu) - 1;
                            }
                            const double prob = (double)nB / (double)(nA + nB);                     // :213
                            const double z2 = rng.uniform_at(a.rng, t, ob + cA + cB, ubase_cur);   // :214
                            if (z2 < prob) { ob += cA; aa += jj; }
                        }
                        --jj;
                    }
                    const int R_sel = aa + 1;
                    double x[EPL];
                    ldv<EPL>(Wst + (size_t)(R_sel - 1) * 2 * d, d, lane, x);   // prev_draw = theta'
                    prev_U = __ldcg(Us + R_sel - 1);                            // = -log pi(theta'), non-finite -> +inf
                    stv<EPL>(Wprev, d, lane, x);
                    good_round = 1;
                    g0_valid = 0;
                }
            }
            n_val += R_n;   // :283
            depth += 1;
            {
                // the far slot just stored is one end; the other end of the whole trajectory comes from the work area
                double xo_[EPL], ro_[EPL];
                ldv<EPL>(dir == 1 ? Wxn : Wxp, d, lane, xo_);
                ldv<EPL>(dir == 1 ? Wrn : Wrp, d, lane, ro_);
#pragma unroll
                for (int k = 0; k < EPL; ++k) xo_[k] = (dir == 1) ? (xf_[k] - xo_[k]) : (xo_[k] - xf_[k]);   // theta+ - theta-
                double dn = lane_dot<EPL, false>(xo_, ro_);   // . r- and . r+ (:286-287; which is which depends on dir, both must be >= 0)
                double dpv = lane_dot<EPL, false>(xo_, rf_);
                warp_sum2<false>(dn, dpv);
                s_val = R_s * ((dn >= 0.0) ? 1 : 0) * ((dpv >= 0.0) ? 1 : 0);   // :289
            }
            __syncwarp();
            ph = PH_DBL_BEGIN;
            break;
        }
        case PH_DRAW_END: {
            if (RNGM == RNG_TAPE) rng.cursor = ubase_cur + ucount;
            // ---- dual averaging (src/nuts.cpp:294-302) ----
            if (t < a.n_adapt) {
                h += (1.0 / ((double)(t + 1) + a.t0)) * (a.delta - (alpha / (double)n_alpha) - h);
                eps = exp(mu - h * sqrt((double)(t + 1)) / a.gamma);
                eps_bar *= exp(pow((double)(t + 1), -a.kappa) * (log(eps) - log(eps_bar)));
            } else {
                eps = eps_bar;
            }
            if (t >= (int)a.n_burnin) {
                double x[EPL];
                ldv<EPL>(Wprev, d, lane, x);
                const long long kept = t - (int)a.n_burnin;
                store_vec<EPL>(a.draws + (chain * a.n_keep + kept) * d, d, lane, x);
                if (a.logp && lane == 0) a.logp[chain * a.n_keep + kept] = -prev_U;
                n_acc += good_round;   // :308
            }
            ++t;
            ph = PH_DRAW_BEGIN;
            break;
        }
        default:
            yielded = true;
            break;
        }
    }
    c.cursor = rng.cursor;
    if (a.dbg && lane == 0) {
        // 0: leaf -> next leaf of the same doubling, 1: a doubling ended, 2: a draw ended, 3: gradient at prev_draw arrived, 4: other
        const int cat = (dbg_ph == PH_LEAF) ? ((c.t != dbg_t) ? 2 : (c.depth != dbg_depth ? 1 : 0)) : (dbg_ph == PH_DBL_G0 ? 3 : 4);
        const unsigned long long dt = (unsigned long long)(clock64() - dbg_t0);
        atomicAdd(a.dbg + 3 * cat, dt);
        atomicAdd(a.dbg + 3 * cat + 1, 1ull);
        atomicMax(a.dbg + 3 * cat + 2, dt);
        if (cat == 0) {   // segments of the common path: finish leapfrog | U-turn tests | walk
            atomicAdd(a.dbg + 15, (unsigned long long)(dbg_t1 - dbg_t0));
            atomicAdd(a.dbg + 16, (unsigned long long)(dbg_t2 - dbg_t1));
            atomicAdd(a.dbg + 17, (unsigned long long)(dbg_t3 - dbg_t2));
        }
    }
}

// One round of every chain's coroutine.
// 128 registers, 4 CTAs per SM at n_dim <= 256 (measured: 96 registers / 5 CTAs spills and is 10 % slower, 168 / 3 is 12 % slower)
template <int EPL, int RNGM, bool MEMO_SH, bool FT> __global__ void __launch_bounds__(LS_WARPS * 32, EPL <= 8 ? 4 : 2) nuts_ls_step(const __grid_constant__ LsArgs a)
{
    __shared__ __align__(16) LsCtl ctl_sh[LS_WARPS];
    __shared__ __align__(16) LsSummary memo_sh[MEMO_SH ? LS_WARPS * LS_TAB_SMEM : 1];
    const int* const lvl_off = a.lvl_off;   // kernel-parameter (constant) bank, indexed dynamically
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chain = a.chain_base + (long long)blockIdx.x * LS_WARPS + warp;
    if (chain >= a.chain_base + a.n_group) return;
    LsCtl* const gctl = a.ctl + chain;
    LsCtl& c = ctl_sh[warp];
    const int d = a.d;
    const int m_max = a.m_max;
    const int memo_n = a.memo_n;
    double* const W = a.work + chain * a.work_stride;
    double* const Wprev = W;
    double* const Wm = W + d;
    double* const Wxp = W + 2 * d;
    double* const Wxn = W + 3 * d;
    double* const Wrp = W + 4 * d;
    double* const Wrn = W + 5 * d;
    double* const TRh = W + 6 * d;
    double* const G0 = W + 7 * d;
    double* const Wst = W + 8 * d;   // state k (1-based): x at Wst + (k-1)*2d, r at + d
    double* const Us = Wst + (size_t)2 * d * m_max;
    double* const Ks = Us + m_max;
    double* const TX = a.TX + chain * d;
    const double* const TY = a.TY + chain * d;
    // summary table: the chain's copy in global memory is its home; with max_tree_depth <= 10 the walk works on a shared-memory
    // copy (loaded here; every update is written through), so that no look-up of the walk pays a global-memory latency
    LsSummary* const memo_g = reinterpret_cast<LsSummary*>(Ks + m_max);
    LsSummary* const memo = MEMO_SH ? memo_sh + warp * LS_TAB_SMEM : memo_g;
    // everything the resume needs is requested at once — control block and summary table straight into shared memory
    // (cp.async: no registers), the three vectors of the pending product into registers — so that one memory latency is exposed
    {
        const char* src = reinterpret_cast<const char*>(gctl);
        char* dst = reinterpret_cast<char*>(&c);
#pragma unroll
        for (int q = 0; q < (LS_CTL_WORDS / 2 + 31) / 32; ++q)
            if (lane + 32 * q < LS_CTL_WORDS / 2) ls_cp_async16(dst + 16 * (lane + 32 * q), src + 16 * (lane + 32 * q));
        if (MEMO_SH) {
            const char* ms = reinterpret_cast<const char*>(memo_g);
            char* md = reinterpret_cast<char*>(memo);
#pragma unroll
            for (int q = 0; q < LS_TAB_SMEM / 32; ++q)
                if (lane + 32 * q < memo_n) ls_cp_async16(md + 16 * (lane + 32 * q), ms + 16 * (lane + 32 * q));
        }
        asm volatile("cp.async.commit_group;\n" ::);
    }
    double xt[EPL], rt[EPL];             // pending position / half-kicked momentum on entry; then the tip of the trajectory
    double yin[EPL];                      // the product the chain asked for
    ldv<EPL, FT>(TX, d, lane, xt);
    ldv<EPL, FT>(TRh, d, lane, rt);
    ldv<EPL, FT>(TY, d, lane, yin);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    if (c.ph == PH_DONE) return;
    ls_resume<EPL, RNGM, MEMO_SH, FT, false>(a, chain, lane, c, memo, memo_g, TX, TRh, xt, rt, yin);
    __syncwarp();
    {
        double2* dst = reinterpret_cast<double2*>(gctl);
        const double2* src = reinterpret_cast<const double2*>(&c);
#pragma unroll
        for (int q = 0; q < (LS_CTL_WORDS / 2 + 31) / 32; ++q)
            if (lane + 32 * q < LS_CTL_WORDS / 2) dst[lane + 32 * q] = src[lane + 32 * q];
    }
}

// ------------------------------------------------------------------------------------------------ persistent variant
// One CTA of 16 warps per SM, one chain per warp, for the whole run: the coroutine's control block, summary table and the
// vectors of the pending leapfrog never leave shared memory, and the round's product is computed by the CTA itself for its
// own 16 chains (mma.sync.m8n8k4.f64, the target's matrix streamed from L2 in 16-row panels by cp.async) — no launches, no
// state reload, no trip through L2 for x / y, and chains only wait for the 15 others of their CTA.  A warp whose chain is
// finished takes the next chain from a global counter, so the 1.73 residency waves of 4096 chains (128 registers per
// thread: 16 chains per SM) cost no idle tail.  Same accumulation order over k as nuts_ls_gemm: identical bits.
// Full tiles only (n_dim = 64, 128, 256), max_tree_depth <= 10.
constexpr int PC_WARPS = 16;   // chains per CTA of the full-size variant; small calls use 8 (template parameter NW)
// NH sub-groups of PC_WARPS / NH warps each run their rounds independently (named barriers): while one sub-group's chains are
// in their tree logic the other's product has the tensor pipe — a round's product and step phases overlap across sub-groups.
template <int EPL, int NH, int NW = PC_WARPS> constexpr size_t pc_smem_bytes()
{
    constexpr int D = 32 * EPL, LDX = D + 4;
    return (size_t)NW * sizeof(LsCtl) + (size_t)NW * LS_TAB_SMEM * sizeof(LsSummary) + (size_t)NW * LDX * 8 + (size_t)2 * NW * D * 8;
}
template <int EPL, int RNGM, int NH, int NW = PC_WARPS> __global__ void __launch_bounds__(NW * 32, 1) nuts_pc_kernel(const __grid_constant__ LsArgs a)
{
    constexpr int D = 32 * EPL, LDX = D + 4, NB = D / 8;
    constexpr int GW = NW / NH;                    // warps (= chains = MMA rows: 8 or 16) per sub-group
    constexpr int MB = GW / 8;                           // 8-row MMA blocks per sub-group
    constexpr int NBW = (NB >= GW) ? NB / GW : 1;        // 8-column output blocks per warp
    extern __shared__ __align__(16) unsigned char pcsm[];
    __shared__ int alive[NW];
    LsCtl* const ctl = reinterpret_cast<LsCtl*>(pcsm);
    LsSummary* const memo_all = reinterpret_cast<LsSummary*>(pcsm + (size_t)NW * sizeof(LsCtl));
    double* const X = reinterpret_cast<double*>(memo_all + (size_t)NW * LS_TAB_SMEM);   // [16][LDX] pending positions (the MMA's A operand)
    double* const RH = X + NW * LDX;                                                      // [16][D]   half-kicked momenta
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp / GW, gw = warp % GW;
    double* const Y = RH + NW * D;                                                        // [16][D]   the products
    const int g = lane >> 2, t4 = lane & 3;
    LsCtl& c = ctl[warp];
    LsSummary* const memo = memo_all + (size_t)warp * LS_TAB_SMEM;
    double* const Xrow = X + warp * LDX;
    double* const RHrow = RH + warp * D;
    const double* const Yrow = Y + warp * D;
    const double* const Xg = X + (size_t)grp * GW * LDX;
    auto gbar = [&]() {
        if (NH == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(GW * 32) : "memory");
    };
    long long chain = -1;
    auto fetch = [&]() {
        int cn = 0;
        if (lane == 0) cn = atomicAdd(a.next_chain, 1);
        cn = __shfl_sync(FULL, cn, 0);
        if (cn < a.n_chains) {
            chain = cn;
            unsigned long long* w = reinterpret_cast<unsigned long long*>(&c);
            for (int i = lane; i < LS_CTL_WORDS; i += 32) w[i] = 0ull;   // ph = PH_INIT0
        } else {
            chain = -1;
        }
        __syncwarp();
    };
    for (int i = lane; i < LDX; i += 32) Xrow[i] = 0.0;
    fetch();
    const bool pw_active = gw * NBW < NB;
    long long pt_step = 0, pt_w1 = 0, pt_prod = 0, pt_w2 = 0, pt_n = 0;
    for (;;) {
        const long long pc0 = a.dbg ? ls_clock() : 0;
        if (chain >= 0) {
            double xt[EPL], rt[EPL], yin[EPL];
#pragma unroll
            for (int m = 0; m < EPL / 2; ++m) {
                const double2 vx = *reinterpret_cast<const double2*>(Xrow + m * 64 + 2 * lane);
                const double2 vr = *reinterpret_cast<const double2*>(RHrow + m * 64 + 2 * lane);
                const double2 vy = *reinterpret_cast<const double2*>(Yrow + m * 64 + 2 * lane);
                xt[2 * m] = vx.x; xt[2 * m + 1] = vx.y; rt[2 * m] = vr.x; rt[2 * m + 1] = vr.y; yin[2 * m] = vy.x; yin[2 * m + 1] = vy.y;
            }
            __syncwarp();
            ls_resume<EPL, RNGM, true, true, true>(a, chain, lane, c, memo, nullptr, Xrow, RHrow, xt, rt, yin);
            __syncwarp();
            if (c.ph == PH_DONE) fetch();   // the new chain's PH_INIT0 runs in the next round (it needs no product)
        }
        if (lane == 0) alive[warp] = chain >= 0 ? 1 : 0;
        const long long pc1 = a.dbg ? ls_clock() : 0;
        gbar();   // every X row of the sub-group is staged, every Y row consumed
        const long long pc2 = a.dbg ? ls_clock() : 0;
        {
            int n = 0;
#pragma unroll
            for (int w = 0; w < GW; ++w) n += alive[grp * GW + w];
            if (n == 0) break;
        }
        // ---- Y[GW][D] = X[GW][D] * A[D][D] ----
        // No shared-memory panels and no barriers inside the product: the B fragment of mma.m8n8k4 (lane (g, t) holds
        // A[4 k4 + t][8 n + g]) is loaded straight from L2 — per instruction four 64-byte row segments, whole sectors — into
        // registers, double-buffered four k4-steps ahead (the warp's registers are free during the product); every warp
        // runs its own column slice at its own pace.  k ascends exactly as in nuts_ls_gemm: identical bits.
        if (pw_active) {
            constexpr int U = 4, NBATCH = D / 4 / U;   // k4-steps per register buffer (8: no faster)
            double acc[MB][NBW][2];
#pragma unroll
            for (int i = 0; i < MB; ++i)
#pragma unroll
                for (int j = 0; j < NBW; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            // With an even number of column blocks per warp, blocks are PAIRED on interleaved columns: block 2p holds the even,
            // block 2p + 1 the odd columns of a 16-column group, so a lane's two B values are adjacent in memory — one 128-bit
            // load per k4-step and pair instead of two 64-bit ones (the LSU pipe was 65 % busy), whole 128-byte lines per
            // warp.  Which column an accumulator belongs to changes, not its arithmetic: the results are the same bits.
            constexpr bool PAIRED = (NBW % 2 == 0);
            const double* const Bcol = a.tdata + (size_t)t4 * D + (size_t)gw * NBW * 8 + (PAIRED ? 2 * g : g);
            double bA[U][NBW], bB[U][NBW];
            auto loadb = [&](double (&b)[U][NBW], int batch) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if constexpr (PAIRED) {
#pragma unroll
                        for (int pr = 0; pr < NBW / 2; ++pr) {
                            const double2 v = __ldg(reinterpret_cast<const double2*>(Bcol + (size_t)(batch * U + u) * 4 * D + pr * 16));
                            b[u][2 * pr] = v.x;
                            b[u][2 * pr + 1] = v.y;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < NBW; ++j) b[u][j] = __ldg(Bcol + (size_t)(batch * U + u) * 4 * D + j * 8);
                    }
                }
            };
            auto compute = [&](const double (&b)[U][NBW], int batch) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double af[MB];
#pragma unroll
                    for (int i = 0; i < MB; ++i) af[i] = Xg[(i * 8 + g) * LDX + (batch * U + u) * 4 + t4];
#pragma unroll
                    for (int j = 0; j < NBW; ++j)
#pragma unroll
                        for (int i = 0; i < MB; ++i) sg_dmma(acc[i][j][0], acc[i][j][1], af[i], b[u][j]);
                }
            };
            loadb(bA, 0);
            for (int batch = 0; batch < NBATCH; batch += 2) {
                loadb(bB, batch + 1);
                compute(bA, batch);
                if (batch + 2 < NBATCH) loadb(bA, batch + 2);
                compute(bB, batch + 1);
            }
            double* const Yg = Y + (size_t)grp * GW * D;
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                if constexpr (PAIRED) {
#pragma unroll
                    for (int pr = 0; pr < NBW / 2; ++pr)
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc)   // accumulator element cc of blocks 2 pr / 2 pr + 1 = columns 2 (2 t + cc), + 1 of the group
                            *reinterpret_cast<double2*>(Yg + (i * 8 + g) * D + gw * NBW * 8 + pr * 16 + 2 * (2 * t4 + cc)) =
                                make_double2(acc[i][2 * pr][cc], acc[i][2 * pr + 1][cc]);
                } else {
#pragma unroll
                    for (int j = 0; j < NBW; ++j)
                        *reinterpret_cast<double2*>(Yg + (i * 8 + g) * D + (gw * NBW + j) * 8 + 2 * t4) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
            }
        }
        const long long pc3 = a.dbg ? ls_clock() : 0;
        gbar();
        if (a.dbg) { pt_step += pc1 - pc0; pt_w1 += pc2 - pc1; pt_prod += pc3 - pc2; pt_w2 += ls_clock() - pc3; ++pt_n; }
    }
    if (a.dbg && lane == 0) {   // per warp: cycles in its resume, waiting for the sub-group, in its part of the product, waiting again
        atomicAdd(a.dbg + 18, (unsigned long long)pt_step); atomicAdd(a.dbg + 19, (unsigned long long)pt_w1);
        atomicAdd(a.dbg + 20, (unsigned long long)pt_prod); atomicAdd(a.dbg + 21, (unsigned long long)pt_w2);
        atomicAdd(a.dbg + 22, (unsigned long long)pt_n);
    }
}

template <int EPL> int ls_launch_step(const LsArgs& a, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((a.n_group + LS_WARPS - 1) / LS_WARPS);
    const bool memo_sh = ls_memo_entries(a.max_depth) <= LS_TAB_SMEM;
    const bool ft = a.d == 32 * EPL;   // full tile: no padding slots, no bounds predicates
    if (a.rng.mode == RNG_PHILOX) {
        if (memo_sh && ft) nuts_ls_step<EPL, RNG_PHILOX, true, true><<<blocks, LS_WARPS * 32, 0, st>>>(a);
        else if (memo_sh) nuts_ls_step<EPL, RNG_PHILOX, true, false><<<blocks, LS_WARPS * 32, 0, st>>>(a);
        else nuts_ls_step<EPL, RNG_PHILOX, false, false><<<blocks, LS_WARPS * 32, 0, st>>>(a);
    } else {
        if (memo_sh) nuts_ls_step<EPL, RNG_TAPE, true, false><<<blocks, LS_WARPS * 32, 0, st>>>(a);
        else nuts_ls_step<EPL, RNG_TAPE, false, false><<<blocks, LS_WARPS * 32, 0, st>>>(a);
    }
    return MCMCB200_OK;
}

}  // namespace

bool nuts_batched_supported(int target_id, int d, bool has_precond, bool strict, bool has_bounds, long long n_chains, int max_depth)
{
    const bool dense_target = target_id == MCMCB200_TARGET_DENSE_GAUSS || target_id == MCMCB200_TARGET_LINREG;
    if (!dense_target || strict || has_bounds || has_precond || (d & 1) || d < 2 || d > 32 * MAX_EPL) return false;
    if (max_depth < 1 || max_depth + 1 > LS_LEVELS || ls_m_max(max_depth) > 255) return false;   // far slot is packed in 8 bits
    if (const char* e = std::getenv("MCMCB200_NUTS_BATCHED")) return e[0] != '0';
    // measured on B200 (C4 target, 400 draws): 512 chains 1.62 s vs 1.72 s on the persistent cooperative kernel, 1024: 1.76 vs 1.74,
    // 4096: 2.95 vs 6.3 — below 512 chains a round's fixed latency (two dependent launches, ~25 us) is all there is
    return n_chains >= 512;
}

long long nuts_batched_work_doubles(long long n_chains, int d, int max_depth)
{
    const long long m = ls_m_max(max_depth);
    const long long jm = ((((long long)(max_depth + 1) * (m + 1) + 1) / 2 + 2) + 1) & ~1ll;        // jmask (unsigned), in doubles, even
    const long long rec = (ls_memo_entries(max_depth) > LS_TAB_SMEM) ? ((n_chains * m + 1) & ~1ll) + ((((n_chains * m + 1) / 2 + 2) + 1) & ~1ll) : 0;   // lalpha_g + ut_g
    return 2ll * RNG_TAB_DOUBLE2 + 2 * n_chains * (long long)d + LS_MAX_GROUPS / 2 + 24 + n_chains * (long long)LS_CTL_WORDS + jm + rec + n_chains * ls_work_per_chain(d, max_depth);
}

int launch_nuts_batched(const NutsLaunch& h, double* work, int* launches, long long* steps_out)
{
    if ((reinterpret_cast<uintptr_t>(h.tdata) | reinterpret_cast<uintptr_t>(work)) & 15) {
        set_error("nuts (chain-batched path): target data must be 16-byte aligned");
        return MCMCB200_ERR_UNSUPPORTED;
    }
    const long long C = h.n_chains;
    const int d = h.d;
    LsArgs a;
    a.n_chains = C; a.d = d; a.target_id = h.target_id; a.tdata = h.tdata; a.x0 = h.x0; a.broadcast_x0 = h.broadcast_x0;
    a.chain_offset = h.chain_offset; a.rng = h.rng; a.draws = h.draws; a.logp = h.logp; a.n_accept = h.n_accept;
    a.step_out = h.step_out; a.n_leapfrog = h.n_leapfrog; a.tape_used = h.tape_used;
    a.n_burnin = h.n_burnin; a.n_keep = h.n_keep; a.n_adapt = h.n_adapt; a.t_end = h.t_end; a.max_depth = h.max_depth;
    a.eps_bar0 = h.eps_bar0; a.delta = h.delta; a.gamma = h.gamma; a.t0 = h.t0; a.kappa = h.kappa;
    double* p = work;
    a.tab = reinterpret_cast<const double2*>(p); p += 2 * RNG_TAB_DOUBLE2;
    a.TX = p; p += C * d;
    a.TY = p; p += C * d;
    a.n_running = reinterpret_cast<int*>(p); p += LS_MAX_GROUPS / 2;
    a.split_dbl_end = (std::getenv("MCMCB200_NUTS_SPLIT") && std::getenv("MCMCB200_NUTS_SPLIT")[0] == '1') ? 1 : 0;   // measured: 3 % slower
    a.dbg = nullptr;
    if (std::getenv("MCMCB200_DEBUG")) {
        a.dbg = reinterpret_cast<unsigned long long*>(p);
        MCMCB200_CUDA_TRY(cudaMemsetAsync(p, 0, 24 * sizeof(double), h.stream));
    }
    p += 24;
    a.ctl = reinterpret_cast<LsCtl*>(p); p += C * LS_CTL_WORDS;
    const int m_max = ls_m_max(h.max_depth);
    a.m_max = m_max;
    a.memo_n = ls_memo_entries(h.max_depth);
    {
        const int Jm = h.max_depth - 1;
        int run = 0;
        for (int j = 0; j < LS_LEVELS; ++j) a.lvl_off[j] = 0;
        for (int j = Jm; j >= 1; --j) { a.lvl_off[j] = run; run += (Jm * (Jm + 1) - j * (j + 1)) / 2 + 1; }
    }
    a.jm_stride = m_max + 1;
    a.jmask = reinterpret_cast<const unsigned*>(p); p += ((((long long)(h.max_depth + 1) * (m_max + 1) + 1) / 2 + 2) + 1) & ~1ll;   // even: 16-byte alignment of what follows
    a.ut_g = nullptr; a.lalpha_g = nullptr;
    if (a.memo_n > LS_TAB_SMEM) {
        a.lalpha_g = p; p += (C * m_max + 1) & ~1ll;
        a.ut_g = reinterpret_cast<unsigned*>(p); p += (((C * m_max + 1) / 2 + 2) + 1) & ~1ll;
    }
    a.work = p;
    {
        // jmask[D][k]: levels j (1 <= j <= D) such that T(j, ao = k - j - 1) is a subtree of a depth-D doubling, i.e. ao is a
        // subset sum of {j + 1, ..., D} (the children of T(j, ao) are T(j - 1, ao) and T(j - 1, ao + j), the root is T(D, 0))
        std::vector<unsigned> jm((size_t)(h.max_depth + 1) * (m_max + 1), 0u);
        for (int D = 1; D <= h.max_depth; ++D) {
            std::vector<char> sums((size_t)m_max + 2, 0);   // subset sums of {j + 1, ..., D}, j descending
            sums[0] = 1;
            for (int j = D; j >= 1; --j) {
                for (int ao = 0; ao + j + 1 <= m_max; ++ao)
                    if (sums[(size_t)ao]) jm[(size_t)D * (m_max + 1) + (ao + j + 1)] |= 1u << j;
                for (int v = m_max + 1 - j; v >= 0; --v)   // sums of {j, ..., D} = sums U (sums + j)
                    if (sums[(size_t)v] && v + j <= m_max + 1) sums[(size_t)(v + j)] = 1;
            }
        }
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(const_cast<unsigned*>(a.jmask), jm.data(), jm.size() * sizeof(unsigned), cudaMemcpyHostToDevice, h.stream));
        MCMCB200_CUDA_TRY(cudaStreamSynchronize(h.stream));   // jm is a local
    }
    a.work_stride = ls_work_per_chain(d, h.max_depth);
    // Persistent variant (nuts_pc_kernel): full tiles up to n_dim = 256 with the summary table in shared memory — one launch for the
    // whole run.  Measured on B200, C4: 2.68 s against 2.96 s for the launched rounds (512 chains: 1.45 s against 1.62 s);
    // MCMCB200_NUTS_PERSIST=0 keeps the launched rounds (tests: the two produce identical bits).
    const bool persist_off = std::getenv("MCMCB200_NUTS_PERSIST") && std::getenv("MCMCB200_NUTS_PERSIST")[0] == '0';
    if (!persist_off && d == 32 * epl_for_dim(d) && epl_for_dim(d) <= 8 && a.memo_n <= LS_TAB_SMEM) {
        cudaStream_t st = h.stream;
        nuts_ls_tables<<<1, 256, 0, st>>>(const_cast<double2*>(a.tab));
        a.next_chain = a.n_running + 1;
        MCMCB200_CUDA_TRY(cudaMemsetAsync(a.n_running, 0, 2 * sizeof(int), st));
        a.chain_base = 0; a.n_group = C;
        int dev = 0, n_sm = 0;
        MCMCB200_CUDA_TRY(cudaGetDevice(&dev));
        MCMCB200_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        // Few chains (a strong-scaling shard: C4 over 8 GPUs = 512 chains): 16 chains per CTA would leave most SMs empty and a
        // CTA's round costs the product of ALL its rows plus its slowest chain — 8 chains per CTA use twice the SMs and halve
        // the product per round (512 chains: 1.45 s -> see DESIGN §4.13).  From 8 x #SMs chains on, 16 per CTA (every B fragment
        // then feeds two MMAs).
        bool small = C <= 8ll * n_sm;
        if (const char* e = std::getenv("MCMCB200_NUTS_PERSIST_NW")) small = (e[0] == '8');
        const int nw = small ? 8 : PC_WARPS;
        const long long want = (C + nw - 1) / nw;
        const long long slots = small ? 2ll * n_sm : n_sm;   // (two 8-chain CTAs fit one SM)
        const unsigned grid = (unsigned)(want < slots ? want : slots);
        const bool philox = a.rng.mode == RNG_PHILOX;
        auto go = [&](auto kern, size_t smem) -> int {
            MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, nw * 32, smem, st>>>(a);
            MCMCB200_CUDA_TRY(cudaGetLastError());
            return MCMCB200_OK;
        };
        int rc = MCMCB200_OK;
        // one group of 16 chains per CTA: every B fragment loaded from L2 feeds two MMAs (two 8-row blocks).  Two independently
        // running groups of 8 (MCMCB200_NUTS_PERSIST_NH=2) overlap product and tree logic but stream the matrix twice: 3.3 s
        const bool halves = std::getenv("MCMCB200_NUTS_PERSIST_NH") && std::getenv("MCMCB200_NUTS_PERSIST_NH")[0] == '2';
#define PC_GO(E) (small  ? (philox ? go(nuts_pc_kernel<E, RNG_PHILOX, 1, 8>, pc_smem_bytes<E, 1, 8>()) : go(nuts_pc_kernel<E, RNG_TAPE, 1, 8>, pc_smem_bytes<E, 1, 8>())) \
                 : halves ? (philox ? go(nuts_pc_kernel<E, RNG_PHILOX, 2>, pc_smem_bytes<E, 2>()) : go(nuts_pc_kernel<E, RNG_TAPE, 2>, pc_smem_bytes<E, 2>())) \
                          : (philox ? go(nuts_pc_kernel<E, RNG_PHILOX, 1>, pc_smem_bytes<E, 1>()) : go(nuts_pc_kernel<E, RNG_TAPE, 1>, pc_smem_bytes<E, 1>())))
        switch (epl_for_dim(d)) {
        case 2: rc = PC_GO(2); break;
        case 4: rc = PC_GO(4); break;
        default: rc = PC_GO(8); break;
        }
#undef PC_GO
        if (rc) return rc;
        if (a.dbg) {
            unsigned long long hd[24];
            MCMCB200_CUDA_TRY(cudaMemcpy(hd, a.dbg, sizeof(hd), cudaMemcpyDeviceToHost));
            if (hd[22]) fprintf(stderr, "nuts (persistent): per warp-round mean cycles: resume %.0f | wait for the group %.0f | product %.0f | wait %.0f  (%llu warp-rounds)\n",
                                (double)hd[18] / hd[22], (double)hd[19] / hd[22], (double)hd[20] / hd[22], (double)hd[21] / hd[22], hd[22]);
            for (int k = 0; k < 5; ++k)
                if (hd[3 * k + 1]) fprintf(stderr, "  resume path %d: count %llu mean %.0f cycles\n", k, hd[3 * k + 1], (double)hd[3 * k] / (double)hd[3 * k + 1]);
            if (hd[1]) fprintf(stderr, "  leaf->leaf segments: finish %.0f | U-turn %.0f | walk %.0f | rest %.0f\n", (double)hd[15] / hd[1], (double)hd[16] / hd[1], (double)hd[17] / hd[1],
                               ((double)hd[0] - hd[15] - hd[16] - hd[17]) / hd[1]);
        }
        *launches = 2;
        if (steps_out) *steps_out = 0;
        return MCMCB200_OK;
    }
    // Chain groups.  A group's rounds (GEMM, step, GEMM, step, ...) are a dependent sequence of short, latency-bound kernels;
    // different groups are independent, so each group runs on its own stream and the GPU overlaps one group's GEMM (fp64
    // tensor pipe) with other groups' step kernels (integer / memory pipes) and fills the SMs a 1024-chain kernel leaves idle.
    // A group's launch sequence is captured once in a CUDA graph of LS_GRAPH_ROUNDS rounds and replayed.
    int G = (int)(C / 512);   // measured on B200, C4 (4096 chains): 2 groups 3.28 s, 4: 3.04, 8: 2.98, 16: 3.01
    if (const char* e = std::getenv("MCMCB200_NUTS_GROUPS")) G = std::atoi(e);
    if (G < 1) G = 1;
    if (G > LS_MAX_GROUPS) G = LS_MAX_GROUPS;
    long long gch = (C + G - 1) / G;
    gch = (gch + 31) / 32 * 32;   // whole GEMM row tiles
    G = (int)((C + gch - 1) / gch);
    struct Res {
        cudaStream_t gs[LS_MAX_GROUPS] = {};
        cudaGraph_t graph[LS_MAX_GROUPS] = {};
        cudaGraphExec_t exec[LS_MAX_GROUPS] = {};
        cudaEvent_t ev_start = nullptr, ev_done[LS_MAX_GROUPS] = {};
        ~Res()
        {
            for (int g = 0; g < LS_MAX_GROUPS; ++g) {
                if (exec[g]) cudaGraphExecDestroy(exec[g]);
                if (graph[g]) cudaGraphDestroy(graph[g]);
                if (ev_done[g]) cudaEventDestroy(ev_done[g]);
                if (gs[g]) cudaStreamDestroy(gs[g]);
            }
            if (ev_start) cudaEventDestroy(ev_start);
        }
    } R;
    cudaStream_t st = h.stream;
    nuts_ls_tables<<<1, 256, 0, st>>>(const_cast<double2*>(a.tab));
    nuts_ls_init<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(a.ctl, C, a.n_running, G, gch);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    MCMCB200_CUDA_TRY(cudaEventCreateWithFlags(&R.ev_start, cudaEventDisableTiming));
    MCMCB200_CUDA_TRY(cudaEventRecord(R.ev_start, st));
    const int epl = epl_for_dim(d);
    LsArgs ga[LS_MAX_GROUPS];
    for (int g = 0; g < G; ++g) {
        ga[g] = a;
        ga[g].chain_base = (long long)g * gch;
        ga[g].n_group = (C - ga[g].chain_base < gch) ? C - ga[g].chain_base : gch;
        ga[g].n_running = a.n_running + g;
        MCMCB200_CUDA_TRY(cudaStreamCreateWithFlags(&R.gs[g], cudaStreamNonBlocking));
        MCMCB200_CUDA_TRY(cudaEventCreateWithFlags(&R.ev_done[g], cudaEventDisableTiming));
        MCMCB200_CUDA_TRY(cudaStreamWaitEvent(R.gs[g], R.ev_start, 0));
    }
    auto step = [&](int g) -> int {
        switch (epl) {
        case 2: return ls_launch_step<2>(ga[g], R.gs[g]);
        case 4: return ls_launch_step<4>(ga[g], R.gs[g]);
        case 8: return ls_launch_step<8>(ga[g], R.gs[g]);
        case 16: return ls_launch_step<16>(ga[g], R.gs[g]);
        default: set_error("nuts (chain-batched path): n_dim=%d unsupported", d); return MCMCB200_ERR_UNSUPPORTED;
        }
    };
    const bool big_tiles = std::getenv("MCMCB200_NUTS_BIG_GEMM") != nullptr;   // the 128 x 64 tiles of mala_wide.cu (same results)
    auto gemm = [&](int g) -> int {
        if (big_tiles) return launch_dgemm_dmma(a.TX + ga[g].chain_base * d, h.tdata, a.TY + ga[g].chain_base * d, ga[g].n_group, d, R.gs[g]);
        const dim3 grid((d + SG_N - 1) / SG_N, (unsigned)((ga[g].n_group + SG_M - 1) / SG_M));
        nuts_ls_gemm<<<grid, SG_THREADS, 0, R.gs[g]>>>(a.TX + ga[g].chain_base * d, h.tdata, a.TY + ga[g].chain_base * d, (int)ga[g].n_group, d);
        return MCMCB200_OK;
    };
    // an upper bound on the rounds a chain can need: per draw and doubling at most m_max new states + one prev_draw gradient,
    // plus the initial step-size search (eps doubles until the energy error exceeds log 2: < 2200 leapfrogs in fp64)
    const long long n_total = h.t_end;
    const long long cap = n_total * (long long)h.max_depth * (ls_m_max(h.max_depth) + 1) + 4096;
    int nl = 2, rc = MCMCB200_OK;
    for (int g = 0; g < G; ++g) {
        if ((rc = step(g))) return rc;   // PH_INIT0: posts the first product request
        if ((rc = gemm(g))) return rc;   // first round issued directly (also sets the GEMM's function attributes outside any capture)
        if ((rc = step(g))) return rc;
        nl += 3;
    }
    MCMCB200_CUDA_TRY(cudaGetLastError());
    bool graphed = !std::getenv("MCMCB200_NO_GRAPH");
    for (int g = 0; g < G && graphed; ++g) {
        if (cudaStreamBeginCapture(R.gs[g], cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); graphed = false; break; }
        for (int r = 0; r < LS_GRAPH_ROUNDS && rc == MCMCB200_OK; ++r) {
            rc = gemm(g);
            if (rc == MCMCB200_OK) rc = step(g);
        }
        const cudaError_t e = cudaStreamEndCapture(R.gs[g], &R.graph[g]);
        if (rc) return rc;
        if (e != cudaSuccess || !R.graph[g] || cudaGraphInstantiate(&R.exec[g], R.graph[g], 0) != cudaSuccess) { cudaGetLastError(); graphed = false; }
    }
    long long rounds = 1;
    int running[LS_MAX_GROUPS];
    bool any = true;
    for (int g = 0; g < G; ++g) running[g] = 1;
    while (any) {
        if (rounds >= cap) { set_error("nuts (chain-batched path): chains still running after %lld rounds", rounds); return MCMCB200_ERR_CUDA; }
        for (int rep = 0; rep < LS_POLL_GRAPHS; ++rep)
            for (int g = 0; g < G; ++g) {
                if (running[g] <= 0) continue;
                if (graphed) {
                    MCMCB200_CUDA_TRY(cudaGraphLaunch(R.exec[g], R.gs[g]));
                } else {
                    for (int r = 0; r < LS_GRAPH_ROUNDS; ++r) {
                        if ((rc = gemm(g))) return rc;
                        if ((rc = step(g))) return rc;
                    }
                }
                nl += 2 * LS_GRAPH_ROUNDS;
            }
        rounds += (long long)LS_POLL_GRAPHS * LS_GRAPH_ROUNDS;
        MCMCB200_CUDA_TRY(cudaGetLastError());
        for (int g = 0; g < G; ++g)
            if (running[g] > 0) MCMCB200_CUDA_TRY(cudaMemcpyAsync(&running[g], a.n_running + g, sizeof(int), cudaMemcpyDeviceToHost, R.gs[g]));
        any = false;
        for (int g = 0; g < G; ++g) {
            MCMCB200_CUDA_TRY(cudaStreamSynchronize(R.gs[g]));
            any = any || running[g] > 0;
        }
    }
    for (int g = 0; g < G; ++g) {
        MCMCB200_CUDA_TRY(cudaEventRecord(R.ev_done[g], R.gs[g]));
        MCMCB200_CUDA_TRY(cudaStreamWaitEvent(st, R.ev_done[g], 0));
    }
    *launches = nl;
    if (a.dbg) {
        unsigned long long hd[18];
        MCMCB200_CUDA_TRY(cudaMemcpy(hd, a.dbg, sizeof(hd), cudaMemcpyDeviceToHost));
        const char* nm[5] = {"leaf->leaf", "doubling end", "draw end", "prev_draw gradient", "other"};
        for (int k = 0; k < 5; ++k)
            if (hd[3 * k + 1]) fprintf(stderr, "  resume path %-20s count %10llu  mean %8.0f cycles  max %8llu\n", nm[k], hd[3 * k + 1], (double)hd[3 * k] / (double)hd[3 * k + 1], hd[3 * k + 2]);
        if (hd[1]) fprintf(stderr, "  leaf->leaf segments (mean cycles): finish leapfrog %.0f | U-turn tests %.0f | walk %.0f | write-back %.0f\n", (double)hd[15] / hd[1],
                           (double)hd[16] / hd[1], (double)hd[17] / hd[1], ((double)hd[0] - hd[15] - hd[16] - hd[17]) / hd[1]);
    }
    if (a.dbg && 0) {}
    if (std::getenv("MCMCB200_DEBUG")) fprintf(stderr, "nuts (chain-batched): %lld rounds, %d group(s) of %lld chains, graphs %s\n", rounds, G, gch, graphed ? "on" : "off");
    if (steps_out) *steps_out = rounds;
    return MCMCB200_OK;
}

}  // namespace mcmcb200
