"""ctypes binding of the C ABI in include/mcmc_b200.h (plumbing for tests and bench.py).

The product is the shared library ``mcmc_b200/libmcmc_b200.so`` (hand-written CUDA for
sm_100a behind an ``extern "C"`` shim) and the C++ drop-in header
``include/mcmc_b200.hpp``; this module only marshals numpy / torch buffers into the
POD structs.  It never falls back to a CPU implementation: if the library is missing
``load()`` raises, and if no CUDA device is usable every ``*_run`` call raises
``McmcB200Error`` with the library's message.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcmc_b200.so")

OK, ERR_INVALID_ARG, ERR_UNKNOWN_TARGET, ERR_UNSUPPORTED, ERR_CUDA, ERR_OOM = range(6)
TARGET_ISO_GAUSS, TARGET_DIAG_GAUSS, TARGET_DENSE_GAUSS, TARGET_LINREG, TARGET_NORMAL_MODEL, TARGET_FUNNEL = range(6)
RNG_PHILOX, RNG_MT19937_TAPE, RNG_USER_TAPE = range(3)
MEM_HOST, MEM_DEVICE = 0, 1
ARITH_FAST, ARITH_STRICT = 0, 1
LAYOUT_CHAIN_ROWS, LAYOUT_COLMAJOR = 0, 1
CHOL_LOWER, CHOL_EIGEN_LLT = 0, 1

c_i32, c_i64, c_u64, c_dbl, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double, ctypes.c_void_p


class Problem(ctypes.Structure):
    _fields_ = [("n_chains", c_i64), ("n_dim", c_i32), ("target_id", c_i32), ("target_data", c_vp),
                ("target_data_len", c_i64), ("initial_vals", c_vp), ("initial_mem", c_i32),
                ("broadcast_initial", c_i32), ("chain_offset", c_i64), ("device", c_i32), ("vals_bound", c_i32),
                ("stream", c_vp), ("lower_bounds", c_vp), ("upper_bounds", c_vp)]


class Rng(ctypes.Structure):
    _fields_ = [("mode", c_i32), ("tape_mem", c_i32), ("seed", c_u64), ("tape", c_vp), ("tape_stride", c_i64)]


class HmcSettings(ctypes.Structure):
    _fields_ = [("n_burnin_draws", c_i64), ("n_keep_draws", c_i64), ("n_leap_steps", c_i64), ("step_size", c_dbl),
                ("precond_mat", c_vp), ("chol_mode", c_i32), ("arith", c_i32)]


class MalaSettings(ctypes.Structure):
    _fields_ = [("n_burnin_draws", c_i64), ("n_keep_draws", c_i64), ("step_size", c_dbl), ("precond_mat", c_vp),
                ("chol_mode", c_i32), ("arith", c_i32)]


class RwmhSettings(ctypes.Structure):
    _fields_ = [("n_burnin_draws", c_i64), ("n_keep_draws", c_i64), ("par_scale", c_dbl), ("cov_mat", c_vp),
                ("chol_mode", c_i32), ("arith", c_i32)]


class NutsSettings(ctypes.Structure):
    _fields_ = [("n_burnin_draws", c_i64), ("n_keep_draws", c_i64), ("n_adapt_draws", c_i64),
                ("target_accept_rate", c_dbl), ("max_tree_depth", c_i64), ("step_size", c_dbl), ("gamma_val", c_dbl),
                ("t0_val", c_dbl), ("kappa_val", c_dbl), ("precond_mat", c_vp), ("chol_mode", c_i32), ("arith", c_i32)]


class RmhmcSettings(ctypes.Structure):
    _fields_ = [("n_burnin_draws", c_i64), ("n_keep_draws", c_i64), ("n_leap_steps", c_i64), ("step_size", c_dbl),
                ("n_fp_steps", c_i64), ("chol_mode", c_i32), ("arith", c_i32), ("metric_id", c_i32), ("reserved0", c_i32)]


class DeSettings(ctypes.Structure):
    _fields_ = [("n_burnin_draws", c_i64), ("n_keep_draws", c_i64), ("n_pop", c_i64), ("jumps", c_i32), ("arith", c_i32),
                ("par_b", c_dbl), ("par_gamma_jump", c_dbl), ("initial_lb", c_vp), ("initial_ub", c_vp)]


class Output(ctypes.Structure):
    _fields_ = [("draws_out", c_vp), ("draws_mem", c_i32), ("draws_layout", c_i32), ("n_accept_draws", c_vp),
                ("logp_out", c_vp), ("step_size_out", c_vp), ("n_leapfrog_out", c_vp), ("kernel_ms", ctypes.c_float),
                ("kernel_launches", c_i32)]


class Summary(ctypes.Structure):
    _fields_ = [("mean", c_vp), ("var", c_vp), ("rhat", c_vp), ("chain_mean", c_vp), ("chain_var", c_vp),
                ("kernel_ms", ctypes.c_float), ("reserved0", c_i32)]


class McmcB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mcmc_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load libmcmc_b200.so; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.mcmcb200_last_error.restype = ctypes.c_char_p
    lib.mcmcb200_target_data_len.restype = c_i64
    lib.mcmcb200_register_target.restype = ctypes.c_int
    for f in ("hmc", "mala", "nuts", "rmhmc", "rwmh", "de"):
        getattr(lib, "mcmcb200_%s_run" % f).restype = ctypes.c_int
    _lib = lib
    return lib


_user_libs = {}


def load_user_library(path):
    """dlopen a USER's target library (built per include/mcmc_b200_device.cuh); its static initialiser registers the
    target with libmcmc_b200.so, after which target_id("<name>") resolves it.  libmcmc_b200.so is loaded first, globally,
    so the user library binds to this very copy."""
    path = os.path.abspath(path)
    if path not in _user_libs:
        ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        load()
        _user_libs[path] = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    return _user_libs[path]


def metric_lookup(name):
    t, m = ctypes.c_int(-1), ctypes.c_int(-1)
    _check(load().mcmcb200_metric_lookup(name.encode(), ctypes.byref(t), ctypes.byref(m)))
    return t.value, m.value


def _check(rc):
    if rc != OK:
        raise McmcB200Error(rc, load().mcmcb200_last_error().decode())


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def target_id(name_or_id):
    if isinstance(name_or_id, str):
        tid = load().mcmcb200_target_lookup(name_or_id.encode())
        if tid < 0:
            raise McmcB200Error(ERR_UNKNOWN_TARGET, "unknown target %r" % name_or_id)
        return tid
    return int(name_or_id)


def _colmajor(mat):
    if mat is None:
        return None
    m = np.asarray(mat, dtype=np.float64)
    return np.ascontiguousarray(m.T)  # bytes = column-major image of m


class _Run:
    """Marshals one run call.  Host (numpy) or device (raw pointer) buffers."""

    def __init__(self, sampler, initial_vals, target, target_data, n_keep, n_burnin, rng_mode, seed, chain_offset,
                 device, stream, tape, draws_out, want_logp, initial_dev_ptr=None, n_chains=None, n_dim=None,
                 draws_dev_ptr=None, lower_bounds=None, upper_bounds=None, layout=0):
        lib = load()
        self.lib = lib
        self.keep = []
        if initial_dev_ptr is not None:
            C, d = int(n_chains), int(n_dim)
            x0_ptr, x0_mem, bcast = c_vp(initial_dev_ptr), MEM_DEVICE, 0
        else:
            x0 = np.ascontiguousarray(initial_vals, dtype=np.float64)
            bcast = 0
            if x0.ndim == 1:
                if n_chains is None:
                    x0 = x0[None, :]
                else:
                    bcast = 1
            C = int(n_chains) if bcast else x0.shape[0]
            d = x0.shape[-1]
            self.keep.append(x0)
            x0_ptr, x0_mem = _np_ptr(x0), MEM_HOST
        td = np.ascontiguousarray(target_data if target_data is not None else [], dtype=np.float64).ravel()
        self.keep.append(td)
        self.C, self.d, self.n_keep = C, d, int(n_keep)
        vb, lo, hi = 0, None, None
        if lower_bounds is not None or upper_bounds is not None:
            lo = np.full(d, -np.inf) if lower_bounds is None else np.ascontiguousarray(lower_bounds, dtype=np.float64)
            hi = np.full(d, np.inf) if upper_bounds is None else np.ascontiguousarray(upper_bounds, dtype=np.float64)
            assert lo.size == d and hi.size == d
            self.keep += [lo, hi]
            vb = 1
        self.problem = Problem(C, d, target_id(target), _np_ptr(td) if td.size else None, td.size, x0_ptr, x0_mem, bcast,
                               int(chain_offset), int(device), vb, c_vp(stream) if stream else None, _np_ptr(lo), _np_ptr(hi))
        tape_ptr, tape_stride, tape_mem = None, 0, MEM_HOST
        if rng_mode == RNG_USER_TAPE:
            tp = np.ascontiguousarray(tape, dtype=np.float64).reshape(C, -1)
            self.keep.append(tp)
            tape_ptr, tape_stride = _np_ptr(tp), tp.shape[1]
        self.rng = Rng(int(rng_mode), tape_mem, int(seed), tape_ptr, tape_stride)
        self.n_accept = np.zeros(C, dtype=np.int64)
        self.step_out = np.zeros(C)
        self.nlf_out = np.zeros(C, dtype=np.int64)
        if draws_dev_ptr is not None:
            self.draws = None
            self.logp = None
            self.out = Output(c_vp(draws_dev_ptr), MEM_DEVICE, int(layout), _np_ptr(self.n_accept), None, _np_ptr(self.step_out),
                              _np_ptr(self.nlf_out), 0.0, 0)
        else:
            self.draws = draws_out if draws_out is not None else np.empty((C, self.n_keep, d))
            assert self.draws.dtype == np.float64 and self.draws.flags.c_contiguous and self.draws.size == C * self.n_keep * d
            self.logp = np.empty((C, self.n_keep)) if want_logp else None
            self.out = Output(_np_ptr(self.draws), MEM_HOST, int(layout), _np_ptr(self.n_accept), _np_ptr(self.logp),
                              _np_ptr(self.step_out), _np_ptr(self.nlf_out), 0.0, 0)

    def result(self):
        r = dict(draws=self.draws, n_accept=self.n_accept, kernel_ms=float(self.out.kernel_ms),
                 kernel_launches=int(self.out.kernel_launches), step_size=self.step_out, n_leapfrog=self.nlf_out)
        if self.logp is not None:
            r["logp"] = self.logp
        return r


_COMMON = dict(target_data=None, n_burnin=1000, n_keep=1000, rng_mode=RNG_PHILOX, seed=0, chain_offset=0, device=-1,
               stream=None, tape=None, draws_out=None, want_logp=False, initial_dev_ptr=None, n_chains=None, n_dim=None,
               draws_dev_ptr=None, lower_bounds=None, upper_bounds=None, layout=0)


def _split(kw):
    common = dict(_COMMON)
    for k in list(kw):
        if k in common:
            common[k] = kw.pop(k)
    return common


def hmc(initial_vals, target, n_leap_steps=1, step_size=1.0, precond_mat=None, chol_mode=CHOL_EIGEN_LLT,
        arith=ARITH_FAST, **kw):
    """Many-chain mcmc::hmc (src/hmc.cpp:233-254).  initial_vals: [n_chains][n_dim]."""
    c = _split(kw)
    assert not kw, kw
    run = _Run("hmc", initial_vals, target, **c)
    pm = _colmajor(precond_mat)
    st = HmcSettings(c["n_burnin"], c["n_keep"], int(n_leap_steps), float(step_size), _np_ptr(pm), chol_mode, arith)
    _check(run.lib.mcmcb200_hmc_run(ctypes.byref(run.problem), ctypes.byref(run.rng), ctypes.byref(st), ctypes.byref(run.out)))
    return run.result()


def mala(initial_vals, target, step_size=1.0, precond_mat=None, chol_mode=CHOL_EIGEN_LLT, arith=ARITH_FAST, **kw):
    """Many-chain mcmc::mala (src/mala.cpp:212-235)."""
    c = _split(kw)
    assert not kw, kw
    run = _Run("mala", initial_vals, target, **c)
    pm = _colmajor(precond_mat)
    st = MalaSettings(c["n_burnin"], c["n_keep"], float(step_size), _np_ptr(pm), chol_mode, arith)
    _check(run.lib.mcmcb200_mala_run(ctypes.byref(run.problem), ctypes.byref(run.rng), ctypes.byref(st), ctypes.byref(run.out)))
    return run.result()


def rwmh(initial_vals, target, par_scale=1.0, cov_mat=None, chol_mode=CHOL_EIGEN_LLT, arith=ARITH_FAST, **kw):
    """Many-chain mcmc::rwmh (src/rwmh.cpp:176-199): proposal x + par_scale * chol(cov_mat) z, value-only target."""
    c = _split(kw)
    assert not kw, kw
    run = _Run("rwmh", initial_vals, target, **c)
    cm = _colmajor(cov_mat)
    st = RwmhSettings(c["n_burnin"], c["n_keep"], float(par_scale), _np_ptr(cm), chol_mode, arith)
    _check(run.lib.mcmcb200_rwmh_run(ctypes.byref(run.problem), ctypes.byref(run.rng), ctypes.byref(st), ctypes.byref(run.out)))
    return run.result()


def nuts(initial_vals, target, step_size=1.0, n_adapt_draws=1000, target_accept_rate=0.55, max_tree_depth=10,
         gamma_val=0.05, t0_val=10.0, kappa_val=0.75, precond_mat=None, chol_mode=CHOL_EIGEN_LLT, arith=ARITH_FAST, **kw):
    """Many-chain mcmc::nuts (src/nuts.cpp:336-359)."""
    c = _split(kw)
    assert not kw, kw
    run = _Run("nuts", initial_vals, target, **c)
    pm = _colmajor(precond_mat)
    st = NutsSettings(c["n_burnin"], c["n_keep"], int(n_adapt_draws), float(target_accept_rate), int(max_tree_depth),
                      float(step_size), float(gamma_val), float(t0_val), float(kappa_val), _np_ptr(pm), chol_mode, arith)
    _check(run.lib.mcmcb200_nuts_run(ctypes.byref(run.problem), ctypes.byref(run.rng), ctypes.byref(st), ctypes.byref(run.out)))
    return run.result()


def rmhmc(initial_vals, target, n_leap_steps=1, step_size=1.0, n_fp_steps=5, chol_mode=CHOL_EIGEN_LLT, arith=ARITH_FAST,
          metric_id=0, **kw):
    """Many-chain mcmc::rmhmc (src/rmhmc.cpp:298-325); the metric is the one registered with the target."""
    c = _split(kw)
    assert not kw, kw
    run = _Run("rmhmc", initial_vals, target, **c)
    st = RmhmcSettings(c["n_burnin"], c["n_keep"], int(n_leap_steps), float(step_size), int(n_fp_steps), chol_mode, arith, int(metric_id), 0)
    _check(run.lib.mcmcb200_rmhmc_run(ctypes.byref(run.problem), ctypes.byref(run.rng), ctypes.byref(st), ctypes.byref(run.out)))
    return run.result()


def de(initial_vals, target, n_pop=100, jumps=False, par_b=1e-4, par_gamma_jump=2.0, initial_lb=None, initial_ub=None,
       arith=ARITH_FAST, **kw):
    """Many-population mcmc::de (src/de.cpp:30-271).  initial_vals: [n_populations][n_dim]; result draws:
    [n_populations][n_keep][n_pop][n_dim] (per population the reference's Cube_t, row-major matrices)."""
    c = _split(kw)
    assert not kw, kw
    x0 = np.ascontiguousarray(initial_vals, dtype=np.float64)
    if x0.ndim == 1:
        x0 = x0[None, :]
    P, d = x0.shape
    n_keep = int(c["n_keep"])
    if c["draws_out"] is None:
        c["draws_out"] = np.empty((P, n_keep * int(n_pop), d))
    c2 = dict(c)
    c2["n_keep"] = n_keep * int(n_pop)   # _Run sizes the output per "kept draw"; a population's kept draw is n_pop rows
    run = _Run("de", x0, target, **c2)
    lo = None if initial_lb is None else np.ascontiguousarray(initial_lb, dtype=np.float64)
    hi = None if initial_ub is None else np.ascontiguousarray(initial_ub, dtype=np.float64)
    st = DeSettings(c["n_burnin"], n_keep, int(n_pop), 1 if jumps else 0, arith, float(par_b), float(par_gamma_jump), _np_ptr(lo), _np_ptr(hi))
    _check(run.lib.mcmcb200_de_run(ctypes.byref(run.problem), ctypes.byref(run.rng), ctypes.byref(st), ctypes.byref(run.out)))
    r = run.result()
    if r["draws"] is not None:
        r["draws"] = r["draws"].reshape(P, n_keep, int(n_pop), d)
    return r


def de_tape(seed, n_pop, n_dim, n_gen, par_b):
    out = np.empty(n_pop * n_dim + n_gen * n_pop * (n_dim + 3))
    _check(load().mcmcb200_de_tape(c_u64(seed), c_i64(n_pop), n_dim, c_i64(n_gen), c_dbl(par_b), _np_ptr(out)))
    return out


def target_eval(target, target_data, x, want_grad=True, arith=ARITH_FAST):
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    n, d = x.shape
    td = np.ascontiguousarray(target_data if target_data is not None else [], dtype=np.float64).ravel()
    val = np.empty(n)
    grad = np.empty((n, d)) if want_grad else None
    _check(lib.mcmcb200_target_eval(target_id(target), _np_ptr(td) if td.size else None, c_i64(td.size), d, c_i64(n),
                                    _np_ptr(x), _np_ptr(val), _np_ptr(grad), arith))
    return val, grad


def mt19937_tape(seed, n_pre_normals, n_draws, n_dim):
    out = np.empty(n_pre_normals + n_draws * (n_dim + 1))
    _check(load().mcmcb200_mt19937_tape(c_u64(seed), c_i64(n_pre_normals), c_i64(n_draws), n_dim, _np_ptr(out)))
    return out


def philox_stream(seed, chain, draw, n_dim, n_unif):
    out = np.empty(n_dim + n_unif)
    _check(load().mcmcb200_philox_stream(c_u64(seed), c_i64(chain), c_i64(draw), n_dim, n_unif, _np_ptr(out)))
    return out


def summarize(draws=None, draws_dev_ptr=None, n_chains=None, n_keep=None, n_dim=None, device=-1, stream=None, per_chain=False):
    """On-device summaries of draws_out[n_chains][n_keep][n_dim] (include/mcmc_b200_summary.h): pooled mean / variance and
    Gelman-Rubin R-hat per element, optionally the per-chain means and variances.  Pass a numpy array (uploaded) or a raw
    device pointer with the three sizes (reduced in place, nothing but the summaries crosses PCIe)."""
    lib = load()
    if draws_dev_ptr is None:
        a = np.ascontiguousarray(draws, dtype=np.float64)
        assert a.ndim == 3
        C, T, d = a.shape
        ptr, mem = _np_ptr(a), MEM_HOST
    else:
        C, T, d = int(n_chains), int(n_keep), int(n_dim)
        ptr, mem = c_vp(draws_dev_ptr), MEM_DEVICE
    mean, var, rhat = np.empty(d), np.empty(d), np.empty(d)
    cm = np.empty((C, d)) if per_chain else None
    cv = np.empty((C, d)) if per_chain else None
    out = Summary(_np_ptr(mean), _np_ptr(var), _np_ptr(rhat), _np_ptr(cm), _np_ptr(cv), 0.0, 0)
    _check(lib.mcmcb200_summarize_draws(ptr, mem, c_i64(C), c_i64(T), d, int(device), c_vp(stream) if stream else None, ctypes.byref(out)))
    r = dict(mean=mean, var=var, rhat=rhat, kernel_ms=float(out.kernel_ms))
    if per_chain:
        r.update(chain_mean=cm, chain_var=cv)
    return r


class Comm:
    """NCCL communicator owned by the library (include/mcmc_b200.h, csrc/gather.cu).  `exchange_id(id_bytes_or_None)` is the
    caller's out-of-band broadcast of rank 0's 128-byte id (bench.py uses torch.distributed for that plumbing)."""

    def __init__(self, world_size, rank, device, exchange_id):
        lib = load()
        buf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            _check(lib.mcmcb200_comm_unique_id(buf, ctypes.c_size_t(128)))
        raw = exchange_id(bytes(buf) if rank == 0 else None)
        idb = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
        self.h = c_vp()
        self.world, self.rank = int(world_size), int(rank)
        _check(lib.mcmcb200_comm_init(idb, ctypes.c_size_t(128), self.world, self.rank, int(device), ctypes.byref(self.h)))

    def allgather_draws(self, local_dev_ptr, chains_per_rank, n_keep, n_dim, full_dev_ptr, stream=None):
        cpr = (c_i64 * self.world)(*[int(c) for c in chains_per_rank])
        _check(load().mcmcb200_allgather_draws(self.h, c_vp(local_dev_ptr), cpr, c_i64(n_keep), int(n_dim), c_vp(full_dev_ptr),
                                               c_vp(stream) if stream else None))

    def destroy(self):
        if self.h:
            load().mcmcb200_comm_destroy(self.h)
            self.h = c_vp()


def fp64_peak(device=-1):
    out = ctypes.c_double(0.0)
    _check(load().mcmcb200_fp64_peak(int(device), ctypes.byref(out)))
    return out.value


def device_count():
    return load().mcmcb200_device_count()


EXPORTED_SYMBOLS = [
    "mcmcb200_hmc_settings_default", "mcmcb200_mala_settings_default", "mcmcb200_nuts_settings_default",
    "mcmcb200_rmhmc_settings_default", "mcmcb200_hmc_run", "mcmcb200_mala_run", "mcmcb200_nuts_run",
    "mcmcb200_rmhmc_run", "mcmcb200_rwmh_settings_default", "mcmcb200_rwmh_run", "mcmcb200_target_lookup", "mcmcb200_target_data_len", "mcmcb200_target_eval",
    "mcmcb200_mt19937_tape", "mcmcb200_philox_stream", "mcmcb200_last_error", "mcmcb200_device_count",
    "mcmcb200_release_workspace", "mcmcb200_summarize_draws", "mcmcb200_de_settings_default", "mcmcb200_de_run", "mcmcb200_de_tape",
    "mcmcb200_metric_lookup", "mcmcb200_register_target", "mcmcb200_allgather_draws", "mcmcb200_comm_unique_id", "mcmcb200_comm_init",
    "mcmcb200_comm_destroy", "mcmcb200_fp64_peak", "mcmcb200_host_alloc", "mcmcb200_host_free",
]
