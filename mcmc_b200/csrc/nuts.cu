// Many-chain NUTS with dual averaging: one persistent kernel, one warp per chain.
//
// Replaces internal::nuts_impl (/root/reference/src/nuts.cpp:30-332), nuts_find_initial_step_size and the
// recursive nuts_build_tree (include/mcmc/nuts.ipp:30-93, 97-241), bug-compatibly.
//
// The recursion is restated per SURVEY Appendix C: a doubling of depth j in direction v starts from
// S = (prev_draw, initial momentum of the draw) and only ever visits the states LF^k S (k leapfrog steps of size
// v*eps): the leaf list of T(j, a) is o_j = o_{j-1} ++ (j + o_{j-1}) offset by a, its "near" slot is offset a+1 and
// its "far" slot is a+j+1 when the first half did not stop (else the first half's far slot).  So instead of the
// reference's 2^j leapfrogs per doubling the kernel computes each distinct state once (at most 1 + j(j+1)/2),
// keeps (x_k, r_k, U_k, K_k) in a per-chain work area and evaluates the merge logic — slice counts, alpha statistics,
// stop flags and U-turn tests — on scalars and on dot products of stored states.
//
// Subtree summaries instead of a replay.  Within one doubling the subtree T(j, a) is a pure function of (j, a) except for
// WHICH leaf its theta' selection picks: (n, s, alpha, n_alpha, far slot) do not depend on the uniforms.  The 2^j - 1
// merges of a depth-j doubling touch only O(j^3) distinct (j, a) pairs (172 of 511 at j = 9), so the kernel memoises the
// summary of every T(j, a) it has built (per warp, tagged with the doubling's epoch) and walks the DAG with an explicit
// stack: each distinct subtree is merged — and U-turn-tested — once per doubling.  The recursion draws its uniforms in
// post-order, one per merge, so a built subtree consumes exactly n_alpha - 1 of them and the uniform of a merge sits at a
// known offset of the draw's stream (Philox is counter-based; the tape is read at cursor + offset).  theta' is therefore
// resolved lazily: only when the doubling's proposal is accepted (src/nuts.cpp:261-267) does the kernel descend from the
// root, drawing the one uniform of each merge on the path (O(j) instead of 2^j - 1 Philox blocks).  Results are those of
// the literal recursion (oracle.cpp build_tree restates it literally; tests compare the two).
//
// Per-chain work area (doubles, dp = n_dim rounded up to even):
//   [0,dp) prev_draw  [dp,2dp) draw momentum  [2dp,3dp) theta+  [3dp,4dp) theta-  [4dp,5dp) r+  [5dp,6dp) r-
//   then for k = 1..m_max: x_k, r_k (2 dp each), then U_k[m_max], K_k[m_max], 8 doubles of saved chain state (segmented
//   runs, see NutsLaunch::t_begin) and — only for max_tree_depth > 10 — the summary table (2 doubles per entry).
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "coop_dmma.cuh"
#include "box.cuh"
#include <math_constants.h>
#include <type_traits>

namespace mcmcb200
{

static __host__ __device__ int nuts_m_max(int max_depth)
{
    const int j = max_depth > 0 ? max_depth - 1 : 0;  // deepest tree built is depth max_depth-1
    return 1 + j * (j + 1) / 2;
}

// Summary table: one entry per (j, a), 1 <= j <= Jm = max_depth - 1, 0 <= a <= a_max(j) = (Jm (Jm + 1) - j (j + 1)) / 2
// (a is a subset sum of {j + 1, ..., Jm}); level j starts at off(j) = sum_{i > j} (a_max(i) + 1).
static __host__ __device__ int nuts_memo_entries(int max_depth)
{
    const int Jm = max_depth - 1;
    int n = 0;
    for (int j = 1; j <= Jm; ++j) n += (Jm * (Jm + 1) - j * (j + 1)) / 2 + 1;
    return n;
}
constexpr int NUTS_TAB_SMEM = 256;   // entries per warp kept in shared memory: covers max_tree_depth <= 10 (249 entries)
constexpr int NUTS_STATE_DOUBLES = 8;

#if (!defined(MCMCB200_TARGET_SLICE) || MCMCB200_TARGET_SLICE == 0) && !defined(MCMCB200_USER_TARGET_TYPE)   // one definition across the per-target translation units (the library's)
long long nuts_work_doubles_per_chain(int d, int max_depth)
{
    const long long dp = (d + 1) & ~1;
    const long long m = nuts_m_max(max_depth);
    const long long e = nuts_memo_entries(max_depth);
    return 6 * dp + 2 * dp * m + 2 * m + 2 + NUTS_STATE_DOUBLES + (e > NUTS_TAB_SMEM ? 2 * e : 0);
}
#endif

constexpr int NUTS_MAX_LEVELS = 22;

struct NutsStack {  // per-warp traversal stack (shared memory): frame = (j, a, phase) + the first half's summary while the second is built
    int j[NUTS_MAX_LEVELS], a[NUTS_MAX_LEVELS], phase[NUTS_MAX_LEVELS], n[NUTS_MAX_LEVELS], nalpha[NUTS_MAX_LEVELS];
    double alpha[NUTS_MAX_LEVELS];
    int lvl_off[NUTS_MAX_LEVELS];   // off(j) of the summary table
};

// summary of a built subtree: alpha statistic + two packed words
//   w0 = n (21 bits) | far slot << 21 (8 bits) | s << 29        w1 = n_alpha (21 bits) | epoch << 21 (11 bits, 0 = empty)
struct NutsSummary { double alpha; unsigned w0, w1; };
constexpr unsigned NUTS_EPOCH_MAX = 2047u;

constexpr int nuts_min_blocks(int epl) { return epl <= 4 ? 4 : (epl == 8 ? 2 : 1); }

// NW = warps (chains) per CTA.  NW = 4 is the independent-warps kernel.  NW = 8 (dense targets, many chains) is the
// CTA-cooperative variant: the 8 chains of a CTA evaluate their gradients in lock-step, so the target's d x d matrix is
// read from L2 once per 8 gradients instead of once per gradient (coop_gemv in warp.cuh; at d = 256 every warp used to
// stream the 512 KB precision matrix by itself and the kernel was L2-bound).  Chains consume different numbers of
// gradients, so a warp whose chain is finished (or that has no chain) keeps serving the remaining chains' products in a
// drain loop until the CTA's active-chain counter reaches zero.  Results are bit-identical to the NW = 4 kernel.
template <class T, int EPL, bool DENSE_M, bool STRICT, int RNGM, bool BOX = false, int NW = WARPS_PER_BLOCK>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 1 : nuts_min_blocks(EPL)) nuts_kernel(const __grid_constant__ NutsLaunch a)
{
    constexpr bool COOP = (NW == 8);
    extern __shared__ double smem[];
    __shared__ double2 rng_tab[RNGM == RNG_PHILOX ? RNG_TAB_DOUBLE2 : 1];
    __shared__ NutsStack stacks[NW];
    __shared__ int n_active;   // COOP: chains of this CTA still running
    __shared__ int coop_want;  // COOP: products requested and not yet served
    __shared__ int coop_seq;   // COOP: rounds run so far
    typedef Ar<STRICT> A;
    if (RNGM == RNG_PHILOX) build_rng_tables(rng_tab);
    // summary tables: the first NW * NUTS_TAB_SMEM * 16 bytes of dynamic shared memory (unused when the table is global)
    NutsSummary* const memo_sh = reinterpret_cast<NutsSummary*>(smem);
    for (int i = threadIdx.x; i < NW * NUTS_TAB_SMEM; i += NW * 32) { memo_sh[i].alpha = 0.0; memo_sh[i].w0 = 0u; memo_sh[i].w1 = 0u; }
    double* const vsm = smem + (size_t)NW * NUTS_TAB_SMEM * 2;   // the vectors / panels follow the tables
    if (COOP && threadIdx.x == 0) {
        const long long left = a.n_chains - (long long)blockIdx.x * NW;
        n_active = left < NW ? (int)left : NW;
        coop_want = 0;
        coop_seq = 0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chain = (long long)blockIdx.x * NW + warp;
    const int d = a.d;
    const int dp = (d + 1) & ~1;
    // FAST arithmetic, dense target, n_dim a multiple of 4 that fits one row per lane: the cooperative product runs as fp64
    // tensor-core MMAs (coop_gemv_body_dmma); its operand layout needs 4 doubles of padding per chain and per matrix column
    const bool coop_tma = COOP && (d % 2 == 0) && d <= 32 * NW && ((reinterpret_cast<uintptr_t>(a.tdata) & 15) == 0);
    const bool coop_dmma = coop_tma && !STRICT && (d % 4 == 0) && a.coop_dmma;
    const int cstride = 2 * dp + (coop_dmma ? 4 : 0);   // doubles between the staged vectors of consecutive chains
    double* tscr = vsm + (size_t)warp * cstride;
    double* mscr = tscr + dp;
    typename std::conditional<COOP, CoopWarpCtx<NW>, WarpCtx>::type w;
    w.lane = lane; w.d = d; w.scr = tscr;
    if constexpr (COOP) {
        // dynamic shared memory: NW x (x, y) vectors, then (when the TMA path applies) two panel buffers and two mbarriers
        w.warp = warp; w.coop_base = vsm; w.coop_stride = cstride; w.phase = 0u; w.n_active = &n_active; w.want = &coop_want;
        w.round_seq = &coop_seq; w.prefetched = 0;
        const bool tma = coop_tma;
        w.dmma = coop_dmma ? 1 : 0;
        w.panel_stride = coop_dmma ? d + 4 : d;
        w.panels = tma ? vsm + (size_t)NW * cstride : nullptr;
        w.mbar = tma ? reinterpret_cast<unsigned long long*>(w.panels + (size_t)2 * COOP_PANEL_COLS * w.panel_stride) : nullptr;
        if (tma) {
            if (threadIdx.x == 0) { mbar_init(w.mbar, 1); mbar_init(w.mbar + 1, 1); mbar_init_fence(); }
            __syncthreads();
        }
    }
    if (chain < a.n_chains) {   // ---- this warp's chain (no early return: COOP warps must reach the drain loop) ----
    NutsStack& st = stacks[warp];
    const int m_max = nuts_m_max(a.max_depth);

    double* W = a.work + chain * a.work_stride;
    double* Wprev = W;
    double* Wm = W + dp;
    double* Wxp = W + 2 * dp;
    double* Wxn = W + 3 * dp;
    double* Wrp = W + 4 * dp;
    double* Wrn = W + 5 * dp;
    double* Wst = W + 6 * dp;  // state k (1-based): x at Wst + (k-1)*2dp, r at + dp
    double* Us = Wst + (size_t)2 * dp * m_max;
    double* Ks = Us + m_max;
    double* Wstate = Ks + m_max + 2;   // saved scalars of a segmented run
    const int memo_n = nuts_memo_entries(a.max_depth);
    NutsSummary* const memo = (memo_n > NUTS_TAB_SMEM) ? reinterpret_cast<NutsSummary*>(Wstate + NUTS_STATE_DOUBLES) : memo_sh + (size_t)warp * NUTS_TAB_SMEM;
    if (memo_n > NUTS_TAB_SMEM)
        for (int i = lane; i < memo_n; i += 32) { memo[i].alpha = 0.0; memo[i].w0 = 0u; memo[i].w1 = 0u; }
    if (lane == 0) {
        const int Jm = a.max_depth - 1;
        int run = 0;
        for (int j = Jm; j >= 1; --j) { st.lvl_off[j] = run; run += (Jm * (Jm + 1) - j * (j + 1)) / 2 + 1; }
        st.lvl_off[0] = 0;
    }
    __syncwarp();

    // K = p.(M^-1 p)/2  (src/nuts.cpp:204, nuts.ipp:140)
    auto kinetic = [&](const double (&p)[EPL]) -> double {
        if (DENSE_M) {
            double t[EPL];
            stage_vec<EPL>(mscr, d, lane, p);
            gemv_cm<EPL, STRICT>(a.Minv_cm, d, lane, mscr, 1.0, t);
            return A::mul(0.5, warp_dot<EPL, STRICT>(p, t));
        }
        return A::mul(0.5, warp_dot<EPL, STRICT>(p, p));
    };
    // p = sqrtM z
    auto momentum = [&](double (&p)[EPL]) {
        if (DENSE_M) {
            double t[EPL];
            stage_vec<EPL>(mscr, d, lane, p);
            gemv_cm<EPL, STRICT>(a.S_cm, d, lane, mscr, 1.0, t);
#pragma unroll
            for (int k = 0; k < EPL; ++k) p[k] = t[k];
        }
    };
    // one leapfrog step of (signed) size e; g holds grad log pi(x) on entry and on exit; returns log pi(new x)
    // (src/nuts.cpp:139-154: half kick, drift with (e M^-1) p, half kick — two gradient calls of which the first
    //  repeats the previous step's last one)
    BoxLane<BOX ? EPL : 1> bx;
    if (BOX) bx.load(a.lb, a.ub, d, lane);
    double Jt[EPL];   // diagonal J(v) belonging to the gradient in gt (dead unless BOX)
    // half kick p + ((e*J)*grad)/2 (src/nuts.cpp:111-126); J only with box constraints
    auto kick = [&](double e, double (&p)[EPL], const double (&g)[EPL]) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            if (BOX) p[k] = A::add(p[k], A::mul(A::mul(Jt[k], A::mul(e, g[k])), 0.5));
            else p[k] = STRICT ? A::add(p[k], A::mul(A::mul(e, g[k]), 0.5)) : fma(0.5 * e, g[k], p[k]);
        }
    };
    auto leapfrog = [&](double e, double (&x)[EPL], double (&p)[EPL], double (&g)[EPL]) -> double {
        kick(e, p, g);
        if (DENSE_M) {
            double t[EPL];
            stage_vec<EPL>(mscr, d, lane, p);
            gemv_cm<EPL, STRICT>(a.Minv_cm, d, lane, mscr, e, t);
#pragma unroll
            for (int k = 0; k < EPL; ++k) x[k] = A::add(x[k], t[k]);
        } else {
#pragma unroll
            for (int k = 0; k < EPL; ++k) x[k] = A::mad(e, p[k], x[k]);
        }
        const double lp = box_eval<T, EPL, STRICT, BOX, true, true, true>(a.tdata, w, bx, x, g, Jt);
        kick(e, p, g);
        return lp;
    };
    double x[EPL], xt[EPL], rt[EPL], gt[EPL];
    // COOP: the leapfrog in two halves around the cooperative product (see the leaf code): begin = first half kick, drift, stage x,
    // post the request; finish = wait until a round has served it, gradient / log pi from the product, second half kick.
    bool pend = false;
    int pend_seq = 0;
    auto begin_next = [&](double e) {
        if constexpr (COOP) {
            kick(e, rt, gt);
#pragma unroll
            for (int k = 0; k < EPL; ++k) xt[k] = A::mad(e, rt[k], xt[k]);
            stage_vec<EPL>(w.scr, d, lane, xt);
            pend_seq = *reinterpret_cast<volatile int*>(&coop_seq);
            if (lane == 0) atomicAdd(&coop_want, 1);
            pend = true;
        }
    };
    auto finish_next = [&](double e) -> double {
        double lp = 0.0;
        if constexpr (COOP) {
            while (__shfl_sync(FULL, *reinterpret_cast<volatile int*>(&coop_seq), 0) == pend_seq) coop_round<STRICT>(a.tdata, w, true);
            w.prefetched = 1;
            lp = box_eval<T, EPL, STRICT, BOX, true, true, true>(a.tdata, w, bx, xt, gt, Jt);
            w.prefetched = 0;
            kick(e, rt, gt);
            pend = false;
        }
        return lp;
    };
    auto cancel_next = [&]() {
        if constexpr (COOP) {
            if (pend) {
                // no round can be running (it needs this warp): either ours was served (nothing to undo) or it is still counted
                if (__shfl_sync(FULL, *reinterpret_cast<volatile int*>(&coop_seq), 0) == pend_seq && lane == 0) atomicSub(&coop_want, 1);
                pend = false;
            }
        }
    };
    auto neg_logp_finite = [](double lp) -> double {
        const double U = -lp;
        return isfinite(U) ? U : CUDART_INF;  // "if (!std::isfinite(prop_U)) prop_U = posinf"
    };

    ChainRng<RNGM> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);
    long long n_lf = 0;
    unsigned epoch = 0;
    double eps, mu, h, eps_bar, prev_U;
    int n_acc = 0;

    if (a.t_begin == 0) {
    load_vec<EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), d, lane, x);
    if (BOX) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) x[k] = bx.transform(BOX ? k : 0, x[k]);   // src/nuts.cpp:158-162
    }
    // ---- pre-loop momentum draw (src/nuts.cpp:166-168) and nuts_find_initial_step_size (nuts.ipp:30-93) ----
    rng.template normals<EPL, false>(a.rng, -1, d, lane, rng_tab, rt);
    momentum(rt);
    eps = 1.0;
    {
        const double pU = neg_logp_finite(box_eval<T, EPL, STRICT, BOX, true, true, true>(a.tdata, w, bx, x, gt, Jt));
        const double pK = kinetic(rt);
#pragma unroll
        for (int k = 0; k < EPL; ++k) xt[k] = x[k];
        double qU = neg_logp_finite(leapfrog(eps, xt, rt, gt));
        ++n_lf;
        double qK = kinetic(rt);
        double dH = A::add(-A::add(qU, qK), A::add(pU, pK));
        int a_val = 2 * (dH > -0.69314718055994530942) - 1;   // > std::log(0.5)
        bool cond = dH > -0.69314718055994530942;            // > -std::log(2)
        while (cond) {
            eps *= (a_val == 1) ? 2.0 : 0.5;                  // step_size *= std::pow(2, a_val)
            qU = neg_logp_finite(leapfrog(eps, xt, rt, gt));   // state is NOT reset between doublings (Q14)
            ++n_lf;
            qK = kinetic(rt);
            dH = A::add(-A::add(qU, qK), A::add(pU, pK));
            a_val = 2 * (dH > -0.69314718055994530942) - 1;
            cond = dH > -0.69314718055994530942;
        }
    }
    mu = log(10.0 * eps);  // src/nuts.cpp:174
    h = 0.0;
    eps_bar = a.eps_bar0;
    prev_U = -box_eval<T, EPL, STRICT, BOX, true, false, true>(a.tdata, w, bx, x, gt, Jt);  // :181 (no finite clamp here)
    store_vec<EPL>(Wprev, d, lane, x);
    } else {
        // a later segment of a run that is driven draw by draw from the host (reference-stream mode): the chain's state
        // was parked in its work area by the previous launch
        load_vec<EPL>(Wprev, d, lane, x);
        eps = Wstate[0]; eps_bar = Wstate[1]; h = Wstate[2]; mu = Wstate[3]; prev_U = Wstate[4];
        n_acc = (int)Wstate[5];
        n_lf = (long long)Wstate[6];
    }

    const int n_burnin = (int)a.n_burnin;
    const int t_end = (int)a.t_end;
    const int kept_before = (int)a.t_begin > n_burnin ? (int)a.t_begin - n_burnin : 0;
    double* out_row = a.draws + (chain * a.n_keep + kept_before) * d;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep + kept_before : nullptr;

    // leaf T(0, ao): state k = ao + 1 of the trajectory (nuts.ipp:132-157)
    struct Leaf { int n, s; };
    auto leaf_ns = [&](int k_state, double log_u) -> Leaf {
        const double Uk = reinterpret_cast<volatile double*>(Us)[k_state - 1], Kk = reinterpret_cast<volatile double*>(Ks)[k_state - 1];
        Leaf l;
        l.n = (log_u <= A::sub(-Uk, Kk)) ? 1 : 0;                    // :146
        l.s = (log_u < A::sub(A::sub(1000.0, Uk), Kk)) ? 1 : 0;      // :147
        return l;
    };

    for (int t = (int)a.t_begin; t < t_end; ++t) {
        int ucount = 0;
        rng.template normals<EPL, false>(a.rng, t, d, lane, rng_tab, rt);   // :200
        const long long ubase_cur = rng.cursor;                             // tape mode: where this draw's uniforms start
        momentum(rt);                                                       // :202
        const double prev_K = kinetic(rt);                                  // :204
        store_vec<EPL>(Wm, d, lane, rt);
        const double log_u = A::sub(A::sub(log(rng.uniform_at(a.rng, t, ucount++, ubase_cur)), prev_U), prev_K);   // :206
        store_vec<EPL>(Wxp, d, lane, x);   // :212-215
        store_vec<EPL>(Wxn, d, lane, x);
        store_vec<EPL>(Wrp, d, lane, rt);
        store_vec<EPL>(Wrn, d, lane, rt);
        __syncwarp();

        int depth = 0, s_val = 1, n_alpha = 0, good_round = 0;
        long long n_val = 1;
        double alpha = 0.0;

        while (s_val == 1 && depth < a.max_depth) {   // :227
            const double zz = rng.uniform_at(a.rng, t, ucount++, ubase_cur);   // :233
            const int dir = (zz <= 0.5) ? -1 : 1;
            const double e_signed = (dir == 1) ? eps : -eps;
            const double H0 = A::add(prev_U, prev_K);

            // tip of the lazily extended trajectory LF^k (prev_draw, draw momentum): always restarts here (Q12)
            int computed = 0;
            // new restart point / direction: the summaries of the previous doubling are void
            if (++epoch > NUTS_EPOCH_MAX) {
                for (int i = lane; i < memo_n; i += 32) memo[i].w1 = 0u;
                __syncwarp();
                epoch = 1;
            }
#pragma unroll
            for (int k = 0; k < EPL; ++k) xt[k] = x[k];
            load_vec<EPL>(Wm, d, lane, rt);
            box_eval<T, EPL, STRICT, BOX, false, true, true>(a.tdata, w, bx, xt, gt, Jt);

            // ---- summary of T(depth, 0): explicit-stack walk over the DAG of distinct subtrees ----
            int R_n = 0, R_s = 0, R_nalpha = 0, R_far = 0;
            double R_alpha = 0.0;
            int level = 0;
            if (lane == 0) { st.j[0] = depth; st.a[0] = 0; st.phase[0] = 0; }
            __syncwarp();
            while (level >= 0) {
                if constexpr (COOP) {
                    // the walk needs no gradients: attend the other chains' product rounds instead of making them wait
                    // — but only once enough requests are pending to share the matrix pass (a.coop_batch, or every chain
                    // that is still running): a round costs the same whether it serves one chain or eight
                    int wv = *reinterpret_cast<volatile int*>(&coop_want);
                    int na = *reinterpret_cast<volatile int*>(&n_active);
                    wv = __shfl_sync(FULL, wv, 0);
                    na = __shfl_sync(FULL, na, 0);
                    if (wv > 0 && wv >= (a.coop_batch < na - 1 ? a.coop_batch : na - 1)) coop_round<STRICT>(a.tdata, w, true);
                }
                const int j = reinterpret_cast<volatile int*>(st.j)[level], ao = reinterpret_cast<volatile int*>(st.a)[level],
                          ph = reinterpret_cast<volatile int*>(st.phase)[level];
                if (j == 0) {
                    const int k_need = ao + 1;
                    while (computed < k_need) {   // extend the trajectory by one leapfrog (nuts.ipp:132)
                        double lp;
                        if constexpr (COOP) {
                            if (!pend) begin_next(e_signed);
                            lp = finish_next(e_signed);
                        } else {
                            lp = leapfrog(e_signed, xt, rt, gt);
                        }
                        ++computed; ++n_lf;
                        const double Uk = neg_logp_finite(lp);   // :134-138
                        const double Kk = kinetic(rt);           // :140
                        store_vec<EPL>(Wst + (size_t)(computed - 1) * 2 * dp, d, lane, xt);
                        store_vec<EPL>(Wst + (size_t)(computed - 1) * 2 * dp + dp, d, lane, rt);
                        if (lane == 0) { Us[computed - 1] = Uk; Ks[computed - 1] = Kk; }
                        __syncwarp();
                        if constexpr (COOP) {
                            // the trajectory continues unless this doubling stops or is complete: start the next state's product NOW,
                            // before the merges and U-turn tests this leaf closes — those then overlap with the other chains' work
                            // instead of delaying the CTA's next round (a state that turns out not to be needed is discarded below)
                            if (a.coop_prefetch && computed < 1 + depth * (depth + 1) / 2) begin_next(e_signed);
                        }
                    }
                    const double Uk = reinterpret_cast<volatile double*>(Us)[k_need - 1], Kk = reinterpret_cast<volatile double*>(Ks)[k_need - 1];
                    const Leaf l = leaf_ns(k_need, log_u);
                    R_n = l.n;
                    R_s = l.s;
                    R_alpha = exp(fmin(0.0, A::add(-A::add(Uk, Kk), H0)));        // :157
                    R_nalpha = 1;
                    R_far = k_need;
                    --level;
                    continue;
                }
                NutsSummary* const ent = memo + st.lvl_off[j] + ao;
                if (ph == 0) {
                    const unsigned w1 = reinterpret_cast<volatile unsigned*>(&ent->w1)[0];
                    if ((w1 >> 21) == epoch) {   // built earlier in this doubling
                        const unsigned w0 = reinterpret_cast<volatile unsigned*>(&ent->w0)[0];
                        R_alpha = reinterpret_cast<volatile double*>(&ent->alpha)[0];
                        R_n = (int)(w0 & 0x1fffffu); R_far = (int)((w0 >> 21) & 0xffu); R_s = (int)((w0 >> 29) & 1u);
                        R_nalpha = (int)(w1 & 0x1fffffu);
                        --level;
                    } else {
                        if (lane == 0) { st.phase[level] = 1; st.j[level + 1] = j - 1; st.a[level + 1] = ao; st.phase[level + 1] = 0; }
                        __syncwarp();
                        ++level;
                    }
                    continue;
                }
                if (ph == 1 && R_s == 1) {   // first half returned in R_* and did not stop: build the second half from far(A)
                    if (lane == 0) {
                        st.n[level] = R_n; st.alpha[level] = R_alpha; st.nalpha[level] = R_nalpha;
                        st.phase[level] = 2;
                        st.j[level + 1] = j - 1; st.a[level + 1] = ao + j; st.phase[level + 1] = 0;
                    }
                    __syncwarp();
                    ++level;
                    continue;
                }
                if (ph == 2) {   // second half returned in R_*
                    // U-turn test on the merged subtree's ends: near = ao+1, far = ao+j+1 (Appendix C).  Its outcome only
                    // matters while the subtree has not stopped (s = s'' * ..., nuts.ipp:229).
                    const int near = ao + 1, far = ao + j + 1;
                    if (R_s == 1) {
                        double xn_[EPL], xf_[EPL], rn_[EPL], rf_[EPL];
                        load_vec<EPL>(Wst + (size_t)(near - 1) * 2 * dp, d, lane, xn_);
                        load_vec<EPL>(Wst + (size_t)(near - 1) * 2 * dp + dp, d, lane, rn_);
                        load_vec<EPL>(Wst + (size_t)(far - 1) * 2 * dp, d, lane, xf_);
                        load_vec<EPL>(Wst + (size_t)(far - 1) * 2 * dp + dp, d, lane, rf_);
                        double diff[EPL];
#pragma unroll
                        for (int k = 0; k < EPL; ++k) diff[k] = (dir == 1) ? A::sub(xf_[k], xn_[k]) : A::sub(xn_[k], xf_[k]);   // pos - neg
                        // dir=+1: pos=far, neg=near; dir=-1: pos=near, neg=far
                        const double d_neg = (dir == 1) ? warp_dot<EPL, STRICT>(diff, rn_) : warp_dot<EPL, STRICT>(diff, rf_);   // :226
                        const double d_pos = (dir == 1) ? warp_dot<EPL, STRICT>(diff, rf_) : warp_dot<EPL, STRICT>(diff, rn_);   // :227
                        R_s = ((d_neg >= 0.0) ? 1 : 0) * ((d_pos >= 0.0) ? 1 : 0);                                                 // :229
                    }
                    R_n = reinterpret_cast<volatile int*>(st.n)[level] + R_n;
                    R_alpha = reinterpret_cast<volatile double*>(st.alpha)[level] + R_alpha;
                    R_nalpha = reinterpret_cast<volatile int*>(st.nalpha)[level] + R_nalpha;
                    R_far = far;
                }
                // (ph == 1 with R_s == 0: the result is the first half's, nuts.ipp:234-239)
                if (lane == 0) {
                    ent->alpha = R_alpha;
                    ent->w0 = (unsigned)R_n | ((unsigned)R_far << 21) | ((unsigned)R_s << 29);
                    ent->w1 = (unsigned)R_nalpha | (epoch << 21);
                }
                __syncwarp();
                --level;
            }
            if constexpr (COOP) cancel_next();   // a prefetched state this doubling did not use
            alpha = R_alpha;   // overwritten by every doubling (Q12)
            n_alpha = R_nalpha;
            const int ubase = ucount;   // the merges of this doubling drew uniforms ubase .. ubase + n_alpha - 2 (post-order)
            ucount += R_nalpha - 1;

            // the far slot of T lands in theta^v / r^v (src/nuts.cpp:241-256)
            {
                double xf_[EPL], rf_[EPL];
                load_vec<EPL>(Wst + (size_t)(R_far - 1) * 2 * dp, d, lane, xf_);
                load_vec<EPL>(Wst + (size_t)(R_far - 1) * 2 * dp + dp, d, lane, rf_);
                store_vec<EPL>(dir == 1 ? Wxp : Wxn, d, lane, xf_);
                store_vec<EPL>(dir == 1 ? Wrp : Wrn, d, lane, rf_);
                __syncwarp();
            }
            if (R_s == 1) {
                const double z3 = rng.uniform_at(a.rng, t, ucount++, ubase_cur);   // :261
                if (z3 < (double)R_n / (double)n_val) {              // :263
                    // theta' of T(depth, 0), resolved now: at every merge on the way down the second half's theta' replaces
                    // the first half's with probability n''/(n' + n'') (nuts.ipp:213-221), the merge's uniform being the one
                    // drawn after both halves were built
                    int jj = depth, aa = 0, ob = ubase;
                    while (jj > 0) {
                        int nA, sA, cA, nB, cB;
                        if (jj == 1) {
                            const Leaf lA = leaf_ns(aa + 1, log_u);
                            nA = lA.n; sA = lA.s; cA = 0;
                        } else {
                            const NutsSummary* eA = memo + st.lvl_off[jj - 1] + aa;
                            const unsigned w0 = reinterpret_cast<const volatile unsigned*>(&eA->w0)[0], w1 = reinterpret_cast<const volatile unsigned*>(&eA->w1)[0];
                            nA = (int)(w0 & 0x1fffffu); sA = (int)((w0 >> 29) & 1u); cA = (int)(w1 & 0x1fffffu) - 1;
                        }
                        if (sA == 1) {
                            if (jj == 1) {
                                nB = leaf_ns(aa + jj + 1, log_u).n; cB = 0;
                            } else {
                                const NutsSummary* eB = memo + st.lvl_off[jj - 1] + aa + jj;
                                nB = (int)(reinterpret_cast<const volatile unsigned*>(&eB->w0)[0] & 0x1fffffu);
                                cB = (int)(reinterpret_cast<const volatile unsigned*>(&eB->w1)[0] & 0x1fffffu) - 1;
                            }
                            const double prob = (double)nB / (double)(nA + nB);                       // :213
                            const double z2 = rng.uniform_at(a.rng, t, ob + cA + cB, ubase_cur);     // :214
                            if (z2 < prob) { ob += cA; aa += jj; }
                        }
                        --jj;
                    }
                    const int R_sel = aa + 1;
                    load_vec<EPL>(Wst + (size_t)(R_sel - 1) * 2 * dp, d, lane, x);   // prev_draw = theta'
                    prev_U = reinterpret_cast<volatile double*>(Us)[R_sel - 1];                                         // = -log pi(theta'), non-finite -> +inf
                    store_vec<EPL>(Wprev, d, lane, x);
                    good_round = 1;
                }
            }
            n_val += R_n;   // :283
            depth += 1;
            {
                double xp_[EPL], xn_[EPL], rp_[EPL], rn_[EPL], diff[EPL];
                load_vec<EPL>(Wxp, d, lane, xp_);
                load_vec<EPL>(Wxn, d, lane, xn_);
                load_vec<EPL>(Wrp, d, lane, rp_);
                load_vec<EPL>(Wrn, d, lane, rn_);
#pragma unroll
                for (int k = 0; k < EPL; ++k) diff[k] = A::sub(xp_[k], xn_[k]);
                const int c1 = warp_dot<EPL, STRICT>(diff, rn_) >= 0.0;   // :286
                const int c2 = warp_dot<EPL, STRICT>(diff, rp_) >= 0.0;   // :287
                s_val = R_s * c1 * c2;                                     // :289
            }
        }
        if (RNGM == RNG_TAPE) rng.cursor = ubase_cur + ucount;

        // ---- dual averaging (src/nuts.cpp:294-302) ----
        if (t < a.n_adapt) {
            h += (1.0 / ((double)(t + 1) + a.t0)) * (a.delta - (alpha / (double)n_alpha) - h);
            eps = exp(mu - h * sqrt((double)(t + 1)) / a.gamma);
            eps_bar *= exp(pow((double)(t + 1), -a.kappa) * (log(eps) - log(eps_bar)));
        } else {
            eps = eps_bar;
        }
        if (t >= n_burnin) {
            if (BOX) {   // src/nuts.cpp:316-323
                double xo[EPL];
#pragma unroll
                for (int k = 0; k < EPL; ++k) xo[k] = bx.inv(BOX ? k : 0, x[k]);
                store_vec<EPL>(out_row, d, lane, xo);
            } else {
                store_vec<EPL>(out_row, d, lane, x);
            }
            out_row += d;
            if (out_lp) {
                if (lane == 0) *out_lp = -prev_U;
                ++out_lp;
            }
            n_acc += good_round;   // :308
        }
    }
    if (lane == 0 && a.save_state) {
        Wstate[0] = eps; Wstate[1] = eps_bar; Wstate[2] = h; Wstate[3] = mu; Wstate[4] = prev_U;
        Wstate[5] = (double)n_acc; Wstate[6] = (double)n_lf;
    }
    if (lane == 0 && a.tape_used) a.tape_used[chain] = rng.cursor;
    if (lane == 0) {
        if (a.n_accept) a.n_accept[chain] = n_acc;
        if (a.step_out) a.step_out[chain] = eps;
        if (a.n_leapfrog) a.n_leapfrog[chain] = n_lf;
        if (COOP) atomicSub(&n_active, 1);
    }
    }   // chain < n_chains
    if constexpr (COOP) {
        // drain: keep serving the cooperative products of the chains that are still running
        __syncwarp();
        while (coop_round<STRICT>(a.tdata, w, true)) {}
    }
}

template <class T, int EPL, bool DENSE_M, bool STRICT, int RNGM, bool BOX = false> static int launch_one(const NutsLaunch& a)
{
    if (a.max_depth + 1 > NUTS_MAX_LEVELS) {
        set_error("nuts: max_tree_depth %d exceeds %d", a.max_depth, NUTS_MAX_LEVELS - 1);
        return MCMCB200_ERR_UNSUPPORTED;
    }
    const int dp = (a.d + 1) & ~1;
    if (T::dense_matrix && !BOX && !DENSE_M && a.coop) {   // dense target, M = I, many chains: cooperative gradients, 8 chains per CTA
        constexpr int NW = (T::dense_matrix && !BOX && !DENSE_M) ? 8 : WARPS_PER_BLOCK;
        const long long blocks = (a.n_chains + NW - 1) / NW;
        const bool tma = (a.d % 2 == 0) && a.d <= 32 * NW && ((reinterpret_cast<uintptr_t>(a.tdata) & 15) == 0);   // same tests as the kernel
        const bool dmma = tma && !STRICT && (a.d % 4 == 0) && a.coop_dmma;
        const size_t smem = (size_t)NW * NUTS_TAB_SMEM * sizeof(NutsSummary) + (size_t)NW * (2 * dp + (dmma ? 4 : 0)) * sizeof(double) +
                            (tma ? (size_t)2 * COOP_PANEL_COLS * (a.d + (dmma ? 4 : 0)) * sizeof(double) + 16 : 0);
        auto kern = nuts_kernel<T, EPL, DENSE_M, STRICT, RNGM, BOX, NW>;
        if (smem > 16 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, NW * 32, smem, a.stream>>>(a);
        MCMCB200_CUDA_TRY(cudaGetLastError());
        return MCMCB200_OK;
    }
    const long long blocks = (a.n_chains + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const size_t smem = (size_t)WARPS_PER_BLOCK * NUTS_TAB_SMEM * sizeof(NutsSummary) +
                        ((T::needs_scratch || DENSE_M) ? (size_t)WARPS_PER_BLOCK * 2 * dp * sizeof(double) : 0);
    auto kern = nuts_kernel<T, EPL, DENSE_M, STRICT, RNGM, BOX>;
    if (smem > 16 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T, int EPL, bool DENSE_M> static int launch_mass(const NutsLaunch& a)
{
    if (a.lb != nullptr) {   // box constraints, with or without a dense mass matrix (src/nuts.cpp:111-126,139-154)
        if (a.rng.mode == RNG_PHILOX)
            return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_PHILOX, true>(a) : launch_one<T, EPL, DENSE_M, false, RNG_PHILOX, true>(a);
        return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_TAPE, true>(a) : launch_one<T, EPL, DENSE_M, false, RNG_TAPE, true>(a);
    }
    if (a.rng.mode == RNG_PHILOX)
        return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_PHILOX>(a) : launch_one<T, EPL, DENSE_M, false, RNG_PHILOX>(a);
    return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_TAPE>(a) : launch_one<T, EPL, DENSE_M, false, RNG_TAPE>(a);
}

template <class T> static int launch_target(const NutsLaunch& a)
{
    const bool dense = a.S_cm != nullptr;
    switch (epl_for_dim(a.d)) {
    MCMCB200_EPL_CASE(2, (dense ? launch_mass<T, 2, true>(a) : launch_mass<T, 2, false>(a)))
    MCMCB200_EPL_CASE(4, (dense ? launch_mass<T, 4, true>(a) : launch_mass<T, 4, false>(a)))
    MCMCB200_EPL_CASE(8, (dense ? launch_mass<T, 8, true>(a) : launch_mass<T, 8, false>(a)))
    default:
        set_error("nuts: n_dim=%d exceeds the register-resident kernels (max %d)", a.d, 256);
        return MCMCB200_ERR_UNSUPPORTED;
    }
}

int MCMCB200_SLICED(launch_nuts)(const NutsLaunch& a)
{
    switch (a.target_id) {
#define X(ID, TYPE) \
    case ID: return launch_target<TYPE>(a);
        MCMCB200_FOREACH_TARGET(X)
#undef X
    default:
        set_error("nuts: unknown target id %d", a.target_id);
        return MCMCB200_ERR_UNKNOWN_TARGET;
    }
}

}  // namespace mcmcb200
