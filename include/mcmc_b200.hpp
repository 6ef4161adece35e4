// mcmc_b200.hpp — C++ drop-in for the hot path of kthohr/mcmc (MCMCLib 2.1.0) on top of the C ABI (mcmc_b200.h).
//
// Keeps the reference's call shape and settings surface:
//
//   reference (include/mcmc/hmc.hpp:43-72)            this header
//   ------------------------------------------------   ------------------------------------------------------------
//   bool mcmc::hmc(const ColVec_t& initial_vals,       bool mcmc::hmc(const ColVec_t& initial_vals,
//        std::function<fp_t(const ColVec_t&,                mcmc::registered_kernel target_log_kernel,
//                           ColVec_t*, void*)> kernel,      Mat_t& draws_out,
//        Mat_t& draws_out, void* target_data                void* target_data /* mcmc::kernel_data* */
//        [, algo_settings_t& settings]);                    [, algo_settings_t& settings]);
//
// and likewise mcmc::mala / mcmc::nuts / mcmc::rmhmc / mcmc::rwmh / mcmc::de (rmhmc's tensor_fn slot takes a registered metric,
// mcmc::device_metric("..."), which must belong to the kernel).  A std::function cannot run on the GPU, so the second argument
// names a REGISTERED __device__ functor (mcmc::device_kernel("iso_gauss"), ...) and `target_data` points to a
// mcmc::kernel_data {values, n} that the library copies to the device.
//
// algo_settings_t and the per-sampler structs have the reference's field names and defaults
// (include/misc/mcmc_structs.hpp:66-134,151-184); n_accept_draws is written back the same way (src/hmc.cpp:220-222).
// draws_out comes back n_keep x n_vals, column-major, exactly like the Eigen/Armadillo Mat_t.
//
// Extras the reference does not have:
//   * many chains per call: pass initial_vals as a Mat_t with one COLUMN per chain and receive a Cube_t
//     (one n_keep x n_vals matrix per chain, the layout of DE's draws_out, src/de.cpp:146,216);
//     chain c uses rng seed  rng_seed_value + c  (the convention of BASELINE.md §3);
//   * settings.b200: RNG mode (MCMCB200_RNG_MT19937_TAPE reproduces the reference's std::mt19937_64 stream and is
//     the default for drop-in parity; MCMCB200_RNG_PHILOX is the in-kernel production generator), arithmetic mode,
//     device ordinal.
//
// Box constraints (vals_bound / lower_bounds / upper_bounds, +-inf = open side) run on the device path for HMC and NUTS (with or
// without a precond_mat), MALA (M = I), RM-HMC, RWMH (with or without a cov_mat) and DE.  Unsupported combinations fail loudly:
// MALA with bounds AND a precond_mat, or bounds on the wide (n_vals > 512) kernels, make the call return false with
// mcmc::last_error() set; nothing is ever silently ignored and nothing is computed on the host.
//
// Vector / matrix types: with MCMC_ENABLE_EIGEN_WRAPPERS or MCMC_ENABLE_ARMA_WRAPPERS defined (as for the reference)
// the Eigen / Armadillo types are used; otherwise a minimal column-major ColVec_t / Mat_t pair is provided.
#ifndef MCMC_B200_HPP
#define MCMC_B200_HPP

#include <cstddef>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "mcmc_b200.h"

#if defined(MCMC_ENABLE_EIGEN_WRAPPERS)
#include <Eigen/Dense>
#elif defined(MCMC_ENABLE_ARMA_WRAPPERS)
#include <armadillo>
#endif

namespace mcmc
{

using uint_t = unsigned int;
// MCMC_FPN_TYPE (include/misc/mcmc_options.hpp:80-99).  double (default): every buffer goes to the device as it is.
// float: the reference's fp32 build — ColVec_t / Mat_t / Cube_t, the settings and target_data are fp32 at this boundary; they
// are widened on the way in, the device computes in fp64 (at least the reference's precision: against a float build of the
// reference the draws agree to fp32 rounding until the reference's own rounding decorrelates the chains) and draws_out is
// narrowed on the way out.  Stated tolerance of the fp32 boundary: per-draw L-inf <= 2^-23 |x| against the fp64 build.
#ifndef MCMC_FPN_TYPE
#define MCMC_FPN_TYPE double
#endif
using fp_t = MCMC_FPN_TYPE;
static_assert(sizeof(fp_t) == sizeof(double) || sizeof(fp_t) == sizeof(float), "MCMC_FPN_TYPE must be double or float");

#if defined(MCMC_ENABLE_EIGEN_WRAPPERS)
using ColVec_t = Eigen::Matrix<fp_t, Eigen::Dynamic, 1>;
using Mat_t = Eigen::Matrix<fp_t, Eigen::Dynamic, Eigen::Dynamic>;
namespace b200_detail
{
inline const fp_t* cdata(const ColVec_t& v) { return v.data(); }
inline const fp_t* cdata(const Mat_t& m) { return m.data(); }
inline fp_t* mdata(Mat_t& m) { return m.data(); }
inline size_t vsize(const ColVec_t& v) { return static_cast<size_t>(v.size()); }
inline size_t msize(const Mat_t& m) { return static_cast<size_t>(m.size()); }
inline size_t mrows(const Mat_t& m) { return static_cast<size_t>(m.rows()); }
inline size_t mcols(const Mat_t& m) { return static_cast<size_t>(m.cols()); }
inline void mresize(Mat_t& m, size_t r, size_t c) { m.resize(static_cast<Eigen::Index>(r), static_cast<Eigen::Index>(c)); }
constexpr int default_chol = MCMCB200_CHOL_EIGEN_LLT;  // BMO_MATOPS_CHOL_LOWER under Eigen (core/cholesky.hpp:37, Q8)
}
#elif defined(MCMC_ENABLE_ARMA_WRAPPERS)
using ColVec_t = arma::Col<fp_t>;
using Mat_t = arma::Mat<fp_t>;
namespace b200_detail
{
inline const fp_t* cdata(const ColVec_t& v) { return v.memptr(); }
inline const fp_t* cdata(const Mat_t& m) { return m.memptr(); }
inline fp_t* mdata(Mat_t& m) { return m.memptr(); }
inline size_t vsize(const ColVec_t& v) { return static_cast<size_t>(v.n_elem); }
inline size_t msize(const Mat_t& m) { return static_cast<size_t>(m.n_elem); }
inline size_t mrows(const Mat_t& m) { return static_cast<size_t>(m.n_rows); }
inline size_t mcols(const Mat_t& m) { return static_cast<size_t>(m.n_cols); }
inline void mresize(Mat_t& m, size_t r, size_t c) { m.set_size(r, c); }
constexpr int default_chol = MCMCB200_CHOL_LOWER;  // arma::chol(A, "lower") (core/cholesky.hpp:31)
}
#else
// minimal stand-ins so the header is usable without a linear-algebra library
class ColVec_t
{
  public:
    ColVec_t() {}
    explicit ColVec_t(size_t n) : d_(n, fp_t(0)) {}
    ColVec_t(std::initializer_list<fp_t> l) : d_(l) {}
    size_t size() const { return d_.size(); }
    void resize(size_t n) { d_.resize(n); }
    fp_t& operator()(size_t i) { return d_[i]; }
    const fp_t& operator()(size_t i) const { return d_[i]; }
    fp_t* data() { return d_.data(); }
    const fp_t* data() const { return d_.data(); }

  private:
    std::vector<fp_t> d_;
};
class Mat_t  // column-major, like Eigen / Armadillo
{
  public:
    Mat_t() : r_(0), c_(0) {}
    Mat_t(size_t r, size_t c) : r_(r), c_(c), d_(r * c, fp_t(0)) {}
    size_t rows() const { return r_; }
    size_t cols() const { return c_; }
    size_t size() const { return d_.size(); }
    void resize(size_t r, size_t c) { r_ = r; c_ = c; d_.resize(r * c); }
    fp_t& operator()(size_t i, size_t j) { return d_[j * r_ + i]; }
    const fp_t& operator()(size_t i, size_t j) const { return d_[j * r_ + i]; }
    fp_t* data() { return d_.data(); }
    const fp_t* data() const { return d_.data(); }

  private:
    size_t r_, c_;
    std::vector<fp_t> d_;
};
namespace b200_detail
{
inline const fp_t* cdata(const ColVec_t& v) { return v.data(); }
inline const fp_t* cdata(const Mat_t& m) { return m.data(); }
inline fp_t* mdata(Mat_t& m) { return m.data(); }
inline size_t vsize(const ColVec_t& v) { return v.size(); }
inline size_t msize(const Mat_t& m) { return m.size(); }
inline size_t mrows(const Mat_t& m) { return m.rows(); }
inline size_t mcols(const Mat_t& m) { return m.cols(); }
inline void mresize(Mat_t& m, size_t r, size_t c) { m.resize(r, c); }
constexpr int default_chol = MCMCB200_CHOL_EIGEN_LLT;
}
#endif

// one matrix per chain (bmo::Cube_t analogue, include/BaseMatrixOps/include/extra/cube_type.hpp:23-68)
class Cube_t
{
  public:
    size_t n_mat() const { return mats_.size(); }
    Mat_t& mat(size_t i) { return mats_[i]; }
    const Mat_t& mat(size_t i) const { return mats_[i]; }
    void set_n_mat(size_t n) { mats_.resize(n); }

  private:
    std::vector<Mat_t> mats_;
};

// ---- settings: field names and defaults of include/misc/mcmc_structs.hpp ------------------------------------
struct hmc_settings_t {
    size_t n_burnin_draws = 1E03;
    size_t n_keep_draws = 1E03;
    int omp_n_threads = -1;  // accepted and ignored: there is no host loop to parallelise
    size_t n_leap_steps = 1;
    fp_t step_size = 1.0;
    Mat_t precond_mat;
    size_t n_accept_draws = 0;  // returned: post-burn-in acceptances (many-chain calls: the mean over chains, rounded, so that
                                // n_accept_draws / n_keep_draws stays the acceptance rate; b200.n_accept_per_chain has every chain)
};
struct nuts_settings_t {
    size_t n_burnin_draws = 1E03;
    size_t n_keep_draws = 1E03;
    int omp_n_threads = -1;
    size_t n_adapt_draws = 1E03;
    fp_t target_accept_rate = 0.55;
    size_t max_tree_depth = size_t(10);
    fp_t step_size = 1.0;  // \bar{\epsilon}_0
    fp_t gamma_val = 0.05;
    fp_t t0_val = 10;
    fp_t kappa_val = 0.75;
    Mat_t precond_mat;
    size_t n_accept_draws = 0;
};
struct rmhmc_settings_t {
    size_t n_burnin_draws = 1E03;
    size_t n_keep_draws = 1E03;
    int omp_n_threads = -1;
    size_t n_leap_steps = 1;
    fp_t step_size = 1.0;
    Mat_t precond_mat;  // never read by the reference either
    size_t n_fp_steps = 5;
    size_t n_accept_draws = 0;
};
struct mala_settings_t {
    size_t n_burnin_draws = 1E03;
    size_t n_keep_draws = 1E03;
    int omp_n_threads = -1;
    fp_t step_size = 1.0;
    Mat_t precond_mat;
    size_t n_accept_draws = 0;
};
struct rwmh_settings_t {   // mcmc_structs.hpp:138-149
    size_t n_burnin_draws = 1E03;
    size_t n_keep_draws = 1E03;
    int omp_n_threads = -1;
    fp_t par_scale = 1.0;
    Mat_t cov_mat;
    size_t n_accept_draws = 0;
};
struct de_settings_t {   // mcmc_structs.hpp:44-62
    bool jumps = false;
    size_t n_pop = 100;
    size_t n_burnin_draws = 1E03;
    size_t n_keep_draws = 1E03;
    int omp_n_threads = -1;
    fp_t par_b = 1E-04;
    fp_t par_gamma = 1.0;        // never read by the reference either (src/de.cpp:58-59 uses 2.38 / sqrt(2 n_vals))
    fp_t par_gamma_jump = 2.0;
    ColVec_t initial_lb;         // defaults to initial_vals - 0.5
    ColVec_t initial_ub;         // defaults to initial_vals + 0.5
    size_t n_accept_draws = 0;
};
struct b200_settings_t {
    int rng_mode = MCMCB200_RNG_MT19937_TAPE;  // reference-compatible stream by default (all samplers, NUTS included)
    int arith = MCMCB200_ARITH_FAST;
    int chol_mode = b200_detail::default_chol;
    int device = -1;
    int rmhmc_metric_id = 0;  // which of the kernel's registered metrics plays tensor_fn in mcmc::rmhmc (0 = its default)
    // many-chain calls: CUDA ordinals to shard the chains over (contiguous shards, one host thread per device, each
    // shard with its global chain offset, so the result does not depend on the list); empty = `device` alone
    std::vector<int> devices;
    std::vector<size_t> n_accept_per_chain;  // returned by many-chain calls
};
struct algo_settings_t {
    size_t rng_seed_value = std::random_device{}();  // mcmc_structs.hpp:155
    bool vals_bound = false;
    ColVec_t lower_bounds;
    ColVec_t upper_bounds;
    hmc_settings_t hmc_settings;
    nuts_settings_t nuts_settings;
    rmhmc_settings_t rmhmc_settings;
    mala_settings_t mala_settings;
    rwmh_settings_t rwmh_settings;
    de_settings_t de_settings;
    b200_settings_t b200;
};

// ---- the registered-functor replacement for the std::function callback --------------------------------------
struct registered_kernel {
    int target_id;
};
struct kernel_data {  // what `void* target_data` points to
    const fp_t* values;
    size_t n;
};
inline registered_kernel device_kernel(const char* name) { return registered_kernel{mcmcb200_target_lookup(name)}; }
// the registered-functor replacement for rmhmc's tensor_fn (include/mcmc/rmhmc.hpp:51): a metric belongs to a log-kernel
// (it reads the same data blob); metric_id < 0 = "the kernel's default metric, or settings.b200.rmhmc_metric_id"
struct registered_metric {
    int target_id;
    int metric_id;
    registered_metric(int t, int m) : target_id(t), metric_id(m) {}
    registered_metric(registered_kernel k) : target_id(k.target_id), metric_id(-1) {}   // old call shape: rmhmc(x0, kernel, kernel, ...)
};
inline registered_metric device_metric(const char* name)
{
    int t = -1, m = -1;
    if (mcmcb200_metric_lookup(name, &t, &m) != MCMCB200_OK) return registered_metric(-1, -1);
    return registered_metric(t, m);
}
namespace b200_detail
{
inline std::string& wrapper_error()
{
    static thread_local std::string e;
    return e;
}
}
// text of the last failure of a mcmc::* call on this thread (the reference has no error channel: it always returns true)
inline const char* last_error()
{
    return b200_detail::wrapper_error().empty() ? mcmcb200_last_error() : b200_detail::wrapper_error().c_str();
}

namespace b200_detail
{

// fp_t <-> device (double) conversions: overloads, so that the double build passes pointers through untouched
inline const double* widen(const double* src, size_t, std::vector<double>&) { return src; }
inline const double* widen(const float* src, size_t n, std::vector<double>& tmp)
{
    if (!src) return nullptr;
    tmp.assign(src, src + n);
    return tmp.data();
}
inline double* device_side(double* direct, size_t, std::vector<double>&) { return direct; }            // results land where they belong
inline double* device_side(float*, size_t n, std::vector<double>& tmp) { tmp.resize(n); return tmp.data(); }   // ... or in a double scratch
inline void narrow(double* dst, const double* src, size_t n) { if (dst != src && n) std::memcpy(dst, src, n * sizeof(double)); }
inline void narrow(float* dst, const double* src, size_t n) { for (size_t i = 0; i < n; ++i) dst[i] = static_cast<float>(src[i]); }

inline void fill_problem(mcmcb200_problem_t& pr, mcmcb200_rng_t& rng, const double* x0, size_t d, size_t n_chains, registered_kernel k,
                         const double* target_values, size_t target_n, const algo_settings_t& s, int rng_mode)
{
    std::memset(&pr, 0, sizeof(pr));
    std::memset(&rng, 0, sizeof(rng));
    pr.n_chains = static_cast<int64_t>(n_chains);
    pr.n_dim = static_cast<int32_t>(d);
    pr.target_id = k.target_id;
    pr.target_data = target_values;
    pr.target_data_len = static_cast<int64_t>(target_n);
    pr.initial_vals = x0;  // d x C column-major == [C][d] chain-major
    pr.initial_mem = MCMCB200_MEM_HOST;
    pr.device = s.b200.device;
    rng.mode = rng_mode;
    rng.seed = s.rng_seed_value;
}

// The library hands draws_out back as [chain][j][t] (MCMCB200_LAYOUT_COLMAJOR: transposed on the device), i.e. every chain's
// block already IS the reference's column-major n_keep x d Mat_t: one chain is copied D2H straight into
// draws_out; many chains arrive in one page-locked staging buffer (the D2H then runs at the PCIe rate) and are copied —
// plain memcpy, chains spread over host threads — into the Cube_t's matrices.
struct pinned_buffer {
    double* p = nullptr;
    size_t n = 0;
    bool pinned = false;
    std::vector<double> fallback;
    explicit pinned_buffer(size_t count) : n(count)
    {
        if (count * sizeof(double) >= (size_t(1) << 24)) p = static_cast<double*>(mcmcb200_host_alloc(count * sizeof(double)));
        pinned = p != nullptr;
        if (!p) { fallback.resize(count); p = fallback.data(); }
    }
    ~pinned_buffer() { if (pinned) mcmcb200_host_free(p); }
    pinned_buffer(const pinned_buffer&) = delete;
    pinned_buffer& operator=(const pinned_buffer&) = delete;
};
inline void unpack(const double* buf, size_t n_chains, size_t n_keep, size_t d, Cube_t& cube)
{
    cube.set_n_mat(n_chains);
    auto work = [&](size_t c_begin, size_t c_end) {
        for (size_t c = c_begin; c < c_end; ++c) {
            mresize(cube.mat(c), n_keep, d);   // allocation and first touch happen on the copying thread
            if (n_keep > 0 && d > 0) narrow(mdata(cube.mat(c)), buf + c * n_keep * d, n_keep * d);
        }
    };
    size_t n_thr = 1;
    if (n_chains > 1 && n_chains * n_keep * d >= (size_t(1) << 22)) {
        n_thr = std::thread::hardware_concurrency();
        if (n_thr == 0) n_thr = 1;
        if (n_thr > 32) n_thr = 32;
        if (n_thr > n_chains) n_thr = n_chains;
    }
    if (n_thr <= 1) { work(0, n_chains); return; }
    std::vector<std::thread> th;
    for (size_t g = 0; g < n_thr; ++g) th.emplace_back(work, n_chains * g / n_thr, n_chains * (g + 1) / n_thr);
    for (auto& t : th) t.join();
}

inline const double* precond_or_null(const Mat_t& m, size_t d, std::vector<double>& tmp)   // src/hmc.cpp:57
{
    return (msize(m) == d * d) ? widen(cdata(m), d * d, tmp) : nullptr;
}

template <class RunFn>
inline bool run(const fp_t* x0, size_t d, size_t n_chains, registered_kernel k, void* target_data, algo_settings_t* sp, size_t n_keep,
                int rng_mode, Mat_t* single, Cube_t* cube, size_t* n_accept_field, RunFn&& fn)
{
    algo_settings_t local;
    algo_settings_t& s = sp ? *sp : local;
    wrapper_error().clear();
    if (s.vals_bound && (vsize(s.lower_bounds) != d || vsize(s.upper_bounds) != d)) {
        // the reference reads n_vals entries of both vectors (determine_bounds_type.hpp:27-57); a short vector is UB there
        wrapper_error() = "mcmc_b200: vals_bound = true needs lower_bounds and upper_bounds of length n_vals";
        return false;
    }
    mcmcb200_problem_t pr;
    mcmcb200_rng_t rng;
    const kernel_data* kd = static_cast<const kernel_data*>(target_data);
    std::vector<double> w_x0, w_td, w_lb, w_ub, w_single;   // used by the fp32 build only
    fill_problem(pr, rng, widen(x0, d * n_chains, w_x0), d, n_chains, k, kd ? widen(kd->values, kd->n, w_td) : nullptr, kd ? kd->n : 0, s, rng_mode);
    if (s.vals_bound) {
        pr.vals_bound = 1;
        pr.lower_bounds = widen(cdata(s.lower_bounds), d, w_lb);
        pr.upper_bounds = widen(cdata(s.upper_bounds), d, w_ub);
    }
    // one chain: the result lands directly in draws_out; many chains: in a page-locked staging buffer
    if (single) mresize(*single, n_keep, d);
    pinned_buffer stage(cube ? n_chains * n_keep * d : 0);
    double* const buf = cube ? stage.p : device_side(mdata(*single), n_keep * d, w_single);
    std::vector<int64_t> acc(n_chains, 0);
    mcmcb200_output_t out;
    std::memset(&out, 0, sizeof(out));
    out.draws_out = buf;
    out.draws_mem = MCMCB200_MEM_HOST;
    out.draws_layout = MCMCB200_LAYOUT_COLMAJOR;
    out.n_accept_draws = acc.data();
    const size_t n_dev = s.b200.devices.size();
    if (n_dev <= 1 || n_chains < 2) {
        if (n_dev == 1) pr.device = s.b200.devices[0];
        if (fn(pr, rng, out, s) != MCMCB200_OK) return false;
    } else {
        // one blocking call drives the listed GPUs (SURVEY §8b "Threading"): the C ABI is re-entrant per host thread
        const size_t n_sh = n_dev < n_chains ? n_dev : n_chains;
        std::vector<int> rcs(n_sh, MCMCB200_OK);
        std::vector<std::string> errs(n_sh);
        std::vector<std::thread> th;
        for (size_t g = 0; g < n_sh; ++g) {
            const size_t c0 = n_chains * g / n_sh, c1 = n_chains * (g + 1) / n_sh;
            th.emplace_back([&, g, c0, c1]() {
                mcmcb200_problem_t p2 = pr;
                mcmcb200_output_t o2 = out;
                p2.n_chains = static_cast<int64_t>(c1 - c0);
                p2.initial_vals = pr.initial_vals + c0 * d;
                p2.chain_offset = pr.chain_offset + static_cast<int64_t>(c0);
                p2.device = s.b200.devices[g];
                o2.draws_out = buf + c0 * n_keep * d;
                o2.n_accept_draws = acc.data() + c0;
                rcs[g] = fn(p2, rng, o2, s);
                if (rcs[g] != MCMCB200_OK) errs[g] = mcmcb200_last_error();   // the C ABI's error text is per thread
            });
        }
        for (auto& t : th) t.join();
        for (size_t g = 0; g < n_sh; ++g)
            if (rcs[g] != MCMCB200_OK) {
                wrapper_error() = errs[g];
                return false;
            }
    }
    if (cube) unpack(buf, n_chains, n_keep, d, *cube);
    else narrow(mdata(*single), buf, n_keep * d);   // (fp64 build: buf IS draws_out)
    if (sp) {  // written back only if a settings object was passed (src/hmc.cpp:220-222)
        long double tot = 0;
        for (size_t c = 0; c < n_chains; ++c) tot += static_cast<long double>(acc[c]);
        *n_accept_field = static_cast<size_t>(tot / static_cast<long double>(n_chains) + 0.5L);   // one chain: its own count
        s.b200.n_accept_per_chain.assign(acc.begin(), acc.end());
    }
    return true;
}

}  // namespace b200_detail

// ================================================= HMC =======================================================
namespace internal
{
inline bool hmc_impl(const fp_t* x0, size_t d, size_t n_chains, registered_kernel k, void* target_data, algo_settings_t* sp, Mat_t* single,
                     Cube_t* cube)
{
    algo_settings_t local;
    algo_settings_t& s = sp ? *sp : local;
    return b200_detail::run(x0, d, n_chains, k, target_data, sp, s.hmc_settings.n_keep_draws, s.b200.rng_mode, single, cube,
                            &s.hmc_settings.n_accept_draws,
                            [&](mcmcb200_problem_t& pr, mcmcb200_rng_t& rng, mcmcb200_output_t& out, algo_settings_t& st) {
                                mcmcb200_hmc_settings_t h;
                                mcmcb200_hmc_settings_default(&h);
                                h.n_burnin_draws = static_cast<int64_t>(st.hmc_settings.n_burnin_draws);
                                h.n_keep_draws = static_cast<int64_t>(st.hmc_settings.n_keep_draws);
                                h.n_leap_steps = static_cast<int64_t>(static_cast<uint_t>(st.hmc_settings.n_leap_steps));  // Q22
                                h.step_size = st.hmc_settings.step_size;
                                std::vector<double> w_pm;
                                h.precond_mat = b200_detail::precond_or_null(st.hmc_settings.precond_mat, d, w_pm);
                                h.chol_mode = st.b200.chol_mode;
                                h.arith = st.b200.arith;
                                return mcmcb200_hmc_run(&pr, &rng, &h, &out);
                            });
}
}  // namespace internal

inline bool hmc(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Mat_t& draws_out, void* target_data)
{
    return internal::hmc_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data, nullptr,
                              &draws_out, nullptr);
}
inline bool hmc(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Mat_t& draws_out, void* target_data,
                algo_settings_t& settings)
{
    return internal::hmc_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data, &settings,
                              &draws_out, nullptr);
}
// many chains: one column of initial_vals per chain
inline bool hmc(const Mat_t& initial_vals, registered_kernel target_log_kernel, Cube_t& draws_out, void* target_data,
                algo_settings_t& settings)
{
    return internal::hmc_impl(b200_detail::cdata(initial_vals), b200_detail::mrows(initial_vals), b200_detail::mcols(initial_vals),
                              target_log_kernel, target_data, &settings, nullptr, &draws_out);
}

// ================================================= MALA ======================================================
namespace internal
{
inline bool mala_impl(const fp_t* x0, size_t d, size_t n_chains, registered_kernel k, void* target_data, algo_settings_t* sp, Mat_t* single,
                      Cube_t* cube)
{
    algo_settings_t local;
    algo_settings_t& s = sp ? *sp : local;
    return b200_detail::run(x0, d, n_chains, k, target_data, sp, s.mala_settings.n_keep_draws, s.b200.rng_mode, single, cube,
                            &s.mala_settings.n_accept_draws,
                            [&](mcmcb200_problem_t& pr, mcmcb200_rng_t& rng, mcmcb200_output_t& out, algo_settings_t& st) {
                                mcmcb200_mala_settings_t m;
                                mcmcb200_mala_settings_default(&m);
                                m.n_burnin_draws = static_cast<int64_t>(st.mala_settings.n_burnin_draws);
                                m.n_keep_draws = static_cast<int64_t>(st.mala_settings.n_keep_draws);
                                m.step_size = st.mala_settings.step_size;
                                std::vector<double> w_pm;
                                m.precond_mat = b200_detail::precond_or_null(st.mala_settings.precond_mat, d, w_pm);
                                m.chol_mode = st.b200.chol_mode;
                                m.arith = st.b200.arith;
                                return mcmcb200_mala_run(&pr, &rng, &m, &out);
                            });
}
}  // namespace internal

inline bool mala(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Mat_t& draws_out, void* target_data)
{
    return internal::mala_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data, nullptr,
                               &draws_out, nullptr);
}
inline bool mala(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Mat_t& draws_out, void* target_data,
                 algo_settings_t& settings)
{
    return internal::mala_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data,
                               &settings, &draws_out, nullptr);
}
inline bool mala(const Mat_t& initial_vals, registered_kernel target_log_kernel, Cube_t& draws_out, void* target_data,
                 algo_settings_t& settings)
{
    return internal::mala_impl(b200_detail::cdata(initial_vals), b200_detail::mrows(initial_vals), b200_detail::mcols(initial_vals),
                               target_log_kernel, target_data, &settings, nullptr, &draws_out);
}

// ================================================= RWMH ======================================================
// bool mcmc::rwmh(initial_vals, target_log_kernel (value only: std::function<fp_t(const ColVec_t&, void*)>), draws_out,
//                 target_data[, settings])   include/mcmc/rwmh.hpp:43-72, src/rwmh.cpp:176-199
namespace internal
{
inline bool rwmh_impl(const fp_t* x0, size_t d, size_t n_chains, registered_kernel k, void* target_data, algo_settings_t* sp, Mat_t* single,
                      Cube_t* cube)
{
    algo_settings_t local;
    algo_settings_t& s = sp ? *sp : local;
    return b200_detail::run(x0, d, n_chains, k, target_data, sp, s.rwmh_settings.n_keep_draws, s.b200.rng_mode, single, cube,
                            &s.rwmh_settings.n_accept_draws,
                            [&](mcmcb200_problem_t& pr, mcmcb200_rng_t& rng, mcmcb200_output_t& out, algo_settings_t& st) {
                                mcmcb200_rwmh_settings_t m;
                                mcmcb200_rwmh_settings_default(&m);
                                m.n_burnin_draws = static_cast<int64_t>(st.rwmh_settings.n_burnin_draws);
                                m.n_keep_draws = static_cast<int64_t>(st.rwmh_settings.n_keep_draws);
                                m.par_scale = st.rwmh_settings.par_scale;
                                std::vector<double> w_pm;
                                m.cov_mat = b200_detail::precond_or_null(st.rwmh_settings.cov_mat, d, w_pm);
                                m.chol_mode = st.b200.chol_mode;
                                m.arith = st.b200.arith;
                                return mcmcb200_rwmh_run(&pr, &rng, &m, &out);
                            });
}
}  // namespace internal

inline bool rwmh(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Mat_t& draws_out, void* target_data)
{
    return internal::rwmh_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data, nullptr,
                               &draws_out, nullptr);
}
inline bool rwmh(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Mat_t& draws_out, void* target_data,
                 algo_settings_t& settings)
{
    return internal::rwmh_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data,
                               &settings, &draws_out, nullptr);
}
inline bool rwmh(const Mat_t& initial_vals, registered_kernel target_log_kernel, Cube_t& draws_out, void* target_data,
                 algo_settings_t& settings)
{
    return internal::rwmh_impl(b200_detail::cdata(initial_vals), b200_detail::mrows(initial_vals), b200_detail::mcols(initial_vals),
                               target_log_kernel, target_data, &settings, nullptr, &draws_out);
}

// ================================================= DE ========================================================
// bool mcmc::de(initial_vals, target_log_kernel (value only), Cube_t& draws_out, target_data[, settings])
// include/mcmc/de.hpp:43-72, src/de.cpp:249-271: draws_out = n_keep matrices of n_pop x n_vals.  Many independent populations
// per call: initial_vals as a d x P matrix (one column per population) and one Cube_t per population.
namespace internal
{
inline bool de_impl(const fp_t* x0, size_t d, size_t n_pops, registered_kernel k, void* target_data, algo_settings_t* sp, Cube_t* cubes)
{
    algo_settings_t local;
    algo_settings_t& s = sp ? *sp : local;
    b200_detail::wrapper_error().clear();
    const de_settings_t& ds = s.de_settings;
    if (s.vals_bound && (b200_detail::vsize(s.lower_bounds) != d || b200_detail::vsize(s.upper_bounds) != d)) {
        b200_detail::wrapper_error() = "mcmc_b200: vals_bound = true needs lower_bounds and upper_bounds of length n_vals";
        return false;
    }
    mcmcb200_problem_t pr;
    mcmcb200_rng_t rng;
    const kernel_data* kd = static_cast<const kernel_data*>(target_data);
    std::vector<double> w_x0, w_td, w_lb, w_ub, w_ilb, w_iub;   // used by the fp32 build only
    b200_detail::fill_problem(pr, rng, b200_detail::widen(x0, d * n_pops, w_x0), d, n_pops, k, kd ? b200_detail::widen(kd->values, kd->n, w_td) : nullptr,
                              kd ? kd->n : 0, s, s.b200.rng_mode);
    if (s.vals_bound) {
        pr.vals_bound = 1;
        pr.lower_bounds = b200_detail::widen(b200_detail::cdata(s.lower_bounds), d, w_lb);
        pr.upper_bounds = b200_detail::widen(b200_detail::cdata(s.upper_bounds), d, w_ub);
    }
    mcmcb200_de_settings_t st;
    mcmcb200_de_settings_default(&st);
    st.n_burnin_draws = static_cast<int64_t>(ds.n_burnin_draws);
    st.n_keep_draws = static_cast<int64_t>(ds.n_keep_draws);
    st.n_pop = static_cast<int64_t>(ds.n_pop);
    st.jumps = ds.jumps ? 1 : 0;
    st.arith = s.b200.arith;
    st.par_b = ds.par_b;
    st.par_gamma_jump = ds.par_gamma_jump;
    st.initial_lb = (b200_detail::vsize(ds.initial_lb) == d) ? b200_detail::widen(b200_detail::cdata(ds.initial_lb), d, w_ilb) : nullptr;   // src/de.cpp:70-71
    st.initial_ub = (b200_detail::vsize(ds.initial_ub) == d) ? b200_detail::widen(b200_detail::cdata(ds.initial_ub), d, w_iub) : nullptr;
    const size_t n_keep = ds.n_keep_draws, n_pop = ds.n_pop;
    std::vector<double> buf(n_pops * n_keep * n_pop * d);
    std::vector<int64_t> acc(n_pops, 0);
    mcmcb200_output_t out;
    std::memset(&out, 0, sizeof(out));
    out.draws_out = buf.data();
    out.draws_mem = MCMCB200_MEM_HOST;
    out.n_accept_draws = acc.data();
    if (mcmcb200_de_run(&pr, &rng, &st, &out) != MCMCB200_OK) return false;
    for (size_t p = 0; p < n_pops; ++p) {   // [generation][member][d] row-major -> n_keep column-major n_pop x d matrices
        cubes[p].set_n_mat(n_keep);
        for (size_t g = 0; g < n_keep; ++g) {
            Mat_t& m = cubes[p].mat(g);
            b200_detail::mresize(m, n_pop, d);
            const double* src = buf.data() + ((p * n_keep + g) * n_pop) * d;
            fp_t* dst = b200_detail::mdata(m);
            for (size_t j = 0; j < d; ++j)
                for (size_t i = 0; i < n_pop; ++i) dst[j * n_pop + i] = static_cast<fp_t>(src[i * d + j]);
        }
    }
    if (sp) s.de_settings.n_accept_draws = static_cast<size_t>(acc[0]);   // src/de.cpp:237-239 (population 0 in many-population calls)
    return true;
}
}  // namespace internal

inline bool de(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Cube_t& draws_out, void* target_data)
{
    return internal::de_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data, nullptr, &draws_out);
}
inline bool de(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Cube_t& draws_out, void* target_data, algo_settings_t& settings)
{
    return internal::de_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data, &settings, &draws_out);
}
// many independent populations: one column of initial_vals and one Cube_t per population
inline bool de(const Mat_t& initial_vals, registered_kernel target_log_kernel, std::vector<Cube_t>& draws_out, void* target_data,
               algo_settings_t& settings)
{
    draws_out.resize(b200_detail::mcols(initial_vals));
    return internal::de_impl(b200_detail::cdata(initial_vals), b200_detail::mrows(initial_vals), b200_detail::mcols(initial_vals), target_log_kernel,
                             target_data, &settings, draws_out.data());
}

// ================================================= NUTS ======================================================
// NUTS consumes a data-dependent number of uniforms per draw, so the reference's mt19937 stream cannot be laid out in
// advance: with rng_mode == MCMCB200_RNG_MT19937_TAPE (the drop-in default) the library drives the kernel draw by draw
// and feeds it the reference stream with a look-ahead pool of uniforms (engine.cu, nuts_run_impl) — same draws as the
// reference, one kernel launch per draw.  MCMCB200_RNG_PHILOX runs the whole chain in one launch.
namespace internal
{
inline bool nuts_impl(const fp_t* x0, size_t d, size_t n_chains, registered_kernel k, void* target_data, algo_settings_t* sp, Mat_t* single,
                      Cube_t* cube)
{
    algo_settings_t local;
    algo_settings_t& s = sp ? *sp : local;
    return b200_detail::run(x0, d, n_chains, k, target_data, sp, s.nuts_settings.n_keep_draws, s.b200.rng_mode, single, cube,
                            &s.nuts_settings.n_accept_draws,
                            [&](mcmcb200_problem_t& pr, mcmcb200_rng_t& rng, mcmcb200_output_t& out, algo_settings_t& st) {
                                mcmcb200_nuts_settings_t n;
                                mcmcb200_nuts_settings_default(&n);
                                n.n_burnin_draws = static_cast<int64_t>(st.nuts_settings.n_burnin_draws);
                                n.n_keep_draws = static_cast<int64_t>(st.nuts_settings.n_keep_draws);
                                n.n_adapt_draws = static_cast<int64_t>(st.nuts_settings.n_adapt_draws);
                                n.target_accept_rate = st.nuts_settings.target_accept_rate;
                                n.max_tree_depth = static_cast<int64_t>(st.nuts_settings.max_tree_depth);
                                n.step_size = st.nuts_settings.step_size;
                                n.gamma_val = st.nuts_settings.gamma_val;
                                n.t0_val = st.nuts_settings.t0_val;
                                n.kappa_val = st.nuts_settings.kappa_val;
                                std::vector<double> w_pm;
                                n.precond_mat = b200_detail::precond_or_null(st.nuts_settings.precond_mat, d, w_pm);
                                n.chol_mode = st.b200.chol_mode;
                                n.arith = st.b200.arith;
                                return mcmcb200_nuts_run(&pr, &rng, &n, &out);
                            });
}
}  // namespace internal

inline bool nuts(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Mat_t& draws_out, void* target_data)
{
    return internal::nuts_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data, nullptr,
                               &draws_out, nullptr);
}
inline bool nuts(const ColVec_t& initial_vals, registered_kernel target_log_kernel, Mat_t& draws_out, void* target_data,
                 algo_settings_t& settings)
{
    return internal::nuts_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, target_data,
                               &settings, &draws_out, nullptr);
}
inline bool nuts(const Mat_t& initial_vals, registered_kernel target_log_kernel, Cube_t& draws_out, void* target_data,
                 algo_settings_t& settings)
{
    return internal::nuts_impl(b200_detail::cdata(initial_vals), b200_detail::mrows(initial_vals), b200_detail::mcols(initial_vals),
                               target_log_kernel, target_data, &settings, nullptr, &draws_out);
}

// ================================================= RM-HMC ====================================================
// The reference takes tensor_fn and tensor_data as separate arguments (include/mcmc/rmhmc.hpp:47-86).  Here tensor_fn
// names a REGISTERED metric (mcmc::device_metric("funnel_softabs"), ...), which must belong to the log-kernel, and it
// reads the log-kernel's data blob: tensor_data must be null or the same pointer as target_data.  Anything else makes the
// call return false with mcmc::last_error() set — nothing is ignored.
namespace internal
{
inline bool rmhmc_impl(const fp_t* x0, size_t d, size_t n_chains, registered_kernel k, registered_metric tensor_fn, void* target_data,
                       void* tensor_data, algo_settings_t* sp, Mat_t* single, Cube_t* cube)
{
    algo_settings_t local;
    algo_settings_t& s = sp ? *sp : local;
    if (tensor_fn.target_id != k.target_id) {
        b200_detail::wrapper_error() = "mcmc_b200: rmhmc tensor_fn must be a metric registered for the same log-kernel (see mcmc::device_metric)";
        return false;
    }
    if (tensor_data != nullptr && tensor_data != target_data) {
        b200_detail::wrapper_error() = "mcmc_b200: rmhmc tensor_data must be null or equal to target_data (registered metrics read the kernel's data blob)";
        return false;
    }
    const int metric_id = tensor_fn.metric_id >= 0 ? tensor_fn.metric_id : s.b200.rmhmc_metric_id;
    return b200_detail::run(x0, d, n_chains, k, target_data, sp, s.rmhmc_settings.n_keep_draws, s.b200.rng_mode, single, cube,
                            &s.rmhmc_settings.n_accept_draws,
                            [&](mcmcb200_problem_t& pr, mcmcb200_rng_t& rng, mcmcb200_output_t& out, algo_settings_t& st) {
                                mcmcb200_rmhmc_settings_t r;
                                mcmcb200_rmhmc_settings_default(&r);
                                r.n_burnin_draws = static_cast<int64_t>(st.rmhmc_settings.n_burnin_draws);
                                r.n_keep_draws = static_cast<int64_t>(st.rmhmc_settings.n_keep_draws);
                                r.n_leap_steps = static_cast<int64_t>(static_cast<uint_t>(st.rmhmc_settings.n_leap_steps));
                                r.step_size = st.rmhmc_settings.step_size;
                                r.n_fp_steps = static_cast<int64_t>(static_cast<uint_t>(st.rmhmc_settings.n_fp_steps));
                                r.chol_mode = st.b200.chol_mode;
                                r.arith = st.b200.arith;
                                r.metric_id = metric_id;
                                return mcmcb200_rmhmc_run(&pr, &rng, &r, &out);
                            });
}
}  // namespace internal

inline bool rmhmc(const ColVec_t& initial_vals, registered_kernel target_log_kernel, registered_metric tensor_fn, Mat_t& draws_out,
                  void* target_data, void* tensor_data)
{
    return internal::rmhmc_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, tensor_fn, target_data,
                                tensor_data, nullptr, &draws_out, nullptr);
}
inline bool rmhmc(const ColVec_t& initial_vals, registered_kernel target_log_kernel, registered_metric tensor_fn, Mat_t& draws_out,
                  void* target_data, void* tensor_data, algo_settings_t& settings)
{
    return internal::rmhmc_impl(b200_detail::cdata(initial_vals), b200_detail::vsize(initial_vals), 1, target_log_kernel, tensor_fn, target_data,
                                tensor_data, &settings, &draws_out, nullptr);
}
inline bool rmhmc(const Mat_t& initial_vals, registered_kernel target_log_kernel, registered_metric tensor_fn, Cube_t& draws_out,
                  void* target_data, void* tensor_data, algo_settings_t& settings)
{
    return internal::rmhmc_impl(b200_detail::cdata(initial_vals), b200_detail::mrows(initial_vals), b200_detail::mcols(initial_vals),
                                target_log_kernel, tensor_fn, target_data, tensor_data, &settings, nullptr, &draws_out);
}

}  // namespace mcmc

#endif  // MCMC_B200_HPP
