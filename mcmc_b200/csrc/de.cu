// Differential-evolution MCMC (ter Braak): many independent POPULATIONS per launch, one warp per population.
//
// Replaces internal::de_impl (/root/reference/src/de.cpp:30-246), SURVEY §8(f) item 4 — the reference's own
// population sampler and the only path with an exchange step: member i's proposal reads two other members' rows,
//   X_prop = X_i + gamma (X_c1 - X_c2) + U(-b, b)^d                    src/de.cpp:166-179
//   accept iff  log pi(X_prop) - log pi(X_i) > temperature * log z      :189-192   (temperature = 1, de.hpp:86-89)
// and the member loop updates X IN PLACE, so member i sees the new rows of members < i (the single-threaded semantics
// of the reference; its OpenMP loop races on exactly that).  A population is therefore sequential in i: one warp walks
// the members, the population matrix X (n_pop x n_dim, row stride dp) lives in the population's work area (L1/L2
// resident: 100 x 128 x 8 B = 100 KB), the members' log-densities in shared memory; parallelism is across populations.
// Every generation after burn-in writes the whole population to draws_out[pop][g][member][:] — the reference's
// Cube_t (n_keep matrices of n_pop x n_vals, src/de.cpp:146,210-212) in row-major order.
//
// Variates.  TAPE: the population's stream in the reference's order — n_pop*d initial uniforms, then per generation and
// member {c1, c2 (indices, stored as doubles, already != i and != each other), d proposal uniforms in (-b, b), z} —
// generated on the host from std::mt19937_64 exactly like the reference does (host_tape.cpp host_de_tape) or supplied
// by the caller.  PHILOX: counter (k, g*n_pop + i + 1, population, 1) — the "uniform #k" stream of rng.cuh — with
// k = 1: c1, 2: c2, 3..d+2: proposal, d+3: z; the initial population uses word 1 = 0 and k = 1 + i*d + j.  Indices come
// from one uniform each (floor(u (n_pop-1)) skipping i; floor(u (n_pop-2)) skipping i and c1): the same distribution as
// the reference's rejection loops.
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "box.cuh"
#include <math_constants.h>

namespace mcmcb200
{

template <class T, int EPL, bool STRICT, int RNGM, bool BOX>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) de_kernel(const __grid_constant__ DeLaunch a)
{
    extern __shared__ double smem[];
    typedef Ar<STRICT> A;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pop = (long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    if (pop >= a.n_chains) return;   // whole warp exits together; no block-level barriers below
    const int d = a.d;
    const int dp = (d + 1) & ~1;
    const int n_pop = a.n_pop;
    double* tv = smem + (size_t)warp * (n_pop + dp);   // log pi of every member
    double* tscr = tv + n_pop;                          // target functor scratch
    const WarpCtx w{lane, d, tscr};
    double* X = a.work + (size_t)pop * (size_t)n_pop * dp;

    BoxLane<BOX ? EPL : 1> bx;
    if (BOX) bx.load(a.lb, a.ub, d, lane);
    double gdummy[EPL], Jdummy[EPL];   // never written (WANT_GRAD = false)

    ChainRng<RNGM> rng;
    rng.init(a.rng, pop, a.chain_offset + pop);
    // uniform #k of "draw" word (Philox) / next tape entry
    auto unif = [&](long long draw, int k) -> double {
        if (RNGM == RNG_PHILOX) return rng.uniform(a.rng, draw, k);
        return rng.uniform(a.rng, 0, 1);   // tape: sequential
    };
    // d uniforms, lane-striped, first one is #k0 (Philox) / the next d tape entries
    auto unif_vec = [&](long long draw, int k0, double (&u)[EPL]) {
        if (RNGM == RNG_PHILOX) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const int j = elem_index(lane, k);
                u[k] = (j < d) ? rng.uniform(a.rng, draw, k0 + j) : 0.0;
            }
        } else {
            if (rng.cursor + d > rng.limit) {
                if (a.rng.err_flag) *a.rng.err_flag = 1;
#pragma unroll
                for (int k = 0; k < EPL; ++k) u[k] = 0.0;
            } else {
#pragma unroll
                for (int k = 0; k < EPL; ++k) {
                    const int j = elem_index(lane, k);
                    u[k] = (j < d) ? rng.tape[rng.cursor + j] : 0.0;
                }
            }
            rng.cursor += d;
        }
    };
    auto value_at = [&](const double (&x)[EPL]) -> double {
        double v = box_eval<T, EPL, STRICT, BOX, true, false, true>(a.tdata, w, bx, x, gdummy, Jdummy);
        return isfinite(v) ? v : -CUDART_INF;   // src/de.cpp:130-132,183-185
    };

    // ---- initial population: X_i = lb + (ub - lb) o U(0,1)^d around initial_vals (src/de.cpp:70-71,118-137) ----
    {
        double lo[EPL], hi[EPL];
        load_vec<EPL>(a.init_lb + (a.init_per_pop ? pop * d : 0), d, lane, lo);
        load_vec<EPL>(a.init_ub + (a.init_per_pop ? pop * d : 0), d, lane, hi);
        for (int i = 0; i < n_pop; ++i) {
            double u[EPL], x[EPL];
            unif_vec(-1, 1 + i * d, u);
#pragma unroll
            for (int k = 0; k < EPL; ++k) x[k] = STRICT ? A::add(lo[k], A::mul(A::sub(hi[k], lo[k]), u[k])) : fma(hi[k] - lo[k], u[k], lo[k]);
            store_vec<EPL>(X + (size_t)i * dp, d, lane, x);
            const double v = value_at(x);
            if (lane == 0) tv[i] = v;
        }
        __syncwarp();
    }

    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep);
    const int n_burnin = (int)a.n_burnin;
    const double b = a.par_b;
    double* out = a.draws + (size_t)pop * (size_t)a.n_keep * (size_t)n_pop * d;

    for (int g = 0; g < n_total; ++g) {
        const double gamma_run = (a.jumps && ((g + 1) % 10 == 0)) ? a.gamma_jump : a.gamma;   // :147-149,214-216
        for (int i = 0; i < n_pop; ++i) {
            const long long word = (long long)g * n_pop + i;
            int c1, c2;
            if (RNGM == RNG_PHILOX) {
                c1 = (int)(unif(word, 1) * (double)(n_pop - 1));
                if (c1 > n_pop - 2) c1 = n_pop - 2;
                if (c1 >= i) ++c1;
                c2 = (int)(unif(word, 2) * (double)(n_pop - 2));
                if (c2 > n_pop - 3) c2 = n_pop - 3;
                const int s0 = i < c1 ? i : c1, s1 = i < c1 ? c1 : i;
                if (c2 >= s0) ++c2;
                if (c2 >= s1) ++c2;
            } else {
                c1 = (int)unif(word, 1);
                c2 = (int)unif(word, 2);
                c1 = c1 < 0 ? 0 : (c1 >= n_pop ? n_pop - 1 : c1);   // a malformed caller tape must not index outside X
                c2 = c2 < 0 ? 0 : (c2 >= n_pop ? n_pop - 1 : c2);
            }
            double r[EPL], xi[EPL], x1[EPL], x2[EPL], prop[EPL];
            unif_vec(word, 3, r);
            if (RNGM == RNG_PHILOX) {   // U(0,1) -> U(-b, b)
#pragma unroll
                for (int k = 0; k < EPL; ++k)
                    r[k] = (elem_index(lane, k) < d) ? (STRICT ? A::sub(A::mul(2.0 * b, r[k]), b) : fma(2.0 * b, r[k], -b)) : 0.0;
            }
            load_vec_rw<EPL>(X + (size_t)i * dp, d, lane, xi);
            load_vec_rw<EPL>(X + (size_t)c1 * dp, d, lane, x1);
            load_vec_rw<EPL>(X + (size_t)c2 * dp, d, lane, x2);
#pragma unroll
            for (int k = 0; k < EPL; ++k)   // (X_i + (X_c1 - X_c2) * gamma) + rand, src/de.cpp:179
                prop[k] = STRICT ? A::add(A::add(xi[k], A::mul(A::sub(x1[k], x2[k]), gamma_run)), r[k]) : fma(x1[k] - x2[k], gamma_run, xi[k]) + r[k];
            const double pv = value_at(prop);
            const double comp = A::sub(pv, reinterpret_cast<volatile double*>(tv)[i]);   // :189
            const double z = unif(word, d + 3);                                           // :190
            if (comp > log(z)) {                                                          // :192 (temperature = 1)
                store_vec<EPL>(X + (size_t)i * dp, d, lane, prop);
                if (lane == 0) tv[i] = pv;
                if (g >= n_burnin) ++n_acc;
            }
            __syncwarp();   // the row and tv[i] are visible to the next member's loads
        }
        if (g >= n_burnin) {   // draws_out.mat(g - n_burnin) = X (:210-212), mapped back with inv_transform when bounded (:223-232)
            double* og = out + (size_t)(g - n_burnin) * (size_t)n_pop * d;
            for (int i = 0; i < n_pop; ++i) {
                double x[EPL];
                load_vec_rw<EPL>(X + (size_t)i * dp, d, lane, x);
                if (BOX) {
#pragma unroll
                    for (int k = 0; k < EPL; ++k) x[k] = bx.inv(BOX ? k : 0, x[k]);
                }
                store_vec<EPL>(og + (size_t)i * d, d, lane, x);
            }
        }
    }
    if (lane == 0 && a.n_accept) a.n_accept[pop] = n_acc;
}

template <class T, int EPL, bool STRICT, int RNGM, bool BOX> static int launch_one(const DeLaunch& a)
{
    const long long blocks = (a.n_chains + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int dp = (a.d + 1) & ~1;
    const size_t smem = (size_t)WARPS_PER_BLOCK * (a.n_pop + dp) * sizeof(double);
    auto kern = de_kernel<T, EPL, STRICT, RNGM, BOX>;
    if (smem > 200 * 1024) { set_error("de: n_pop=%d does not fit the per-warp shared-memory table", a.n_pop); return MCMCB200_ERR_UNSUPPORTED; }
    if (smem > 16 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T, int EPL> static int launch_epl(const DeLaunch& a)
{
    const bool box = a.lb != nullptr;
#define DE_MODE(S, R) (box ? launch_one<T, EPL, S, R, true>(a) : launch_one<T, EPL, S, R, false>(a))
    if (a.rng.mode == RNG_PHILOX) return a.strict ? DE_MODE(true, RNG_PHILOX) : DE_MODE(false, RNG_PHILOX);
    return a.strict ? DE_MODE(true, RNG_TAPE) : DE_MODE(false, RNG_TAPE);
#undef DE_MODE
}

template <class T> static int launch_target(const DeLaunch& a)
{
    switch (epl_for_dim(a.d)) {
    MCMCB200_EPL_CASE(2, (launch_epl<T, 2>(a)))
    MCMCB200_EPL_CASE(4, (launch_epl<T, 4>(a)))
    MCMCB200_EPL_CASE(8, (launch_epl<T, 8>(a)))
    MCMCB200_EPL_CASE(16, (launch_epl<T, 16>(a)))
    default:
        set_error("de: n_dim=%d exceeds the register-resident kernels (max %d)", a.d, 32 * MAX_EPL);
        return MCMCB200_ERR_UNSUPPORTED;
    }
}

int MCMCB200_SLICED(launch_de)(const DeLaunch& a)
{
    switch (a.target_id) {
#define X(ID, TYPE) \
    case ID: return launch_target<TYPE>(a);
        MCMCB200_FOREACH_TARGET(X)
#undef X
    default:
        set_error("de: unknown target id %d", a.target_id);
        return MCMCB200_ERR_UNKNOWN_TARGET;
    }
}

}  // namespace mcmcb200
