"""ctypes bindings for the TEST-ONLY CPU checkers under oracle/.

* ``Oracle``    — oracle/liboracle.so, the restated samplers (oracle/oracle.cpp).
* ``Reference`` — oracle/_ref/libmcmc_ref_{strict,fast}.so, the UNMODIFIED reference
  sources behind oracle/ref_driver.cpp (present only where oracle/Makefile built
  them; the GPU box gets the prebuilt files through the gpurun snapshot).

Nothing under mcmc_b200/ imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

HMC, MALA, NUTS, RMHMC, RWMH = 0, 1, 2, 3, 4
RNG_MT, RNG_TAPE, RNG_PHILOX = 0, 1, 2
SUM_SEQ, SUM_WARP = 0, 1
TGT_ISO_GAUSS, TGT_DIAG_GAUSS, TGT_DENSE_GAUSS, TGT_LINREG, TGT_NORMAL_MODEL, TGT_FUNNEL = 0, 1, 2, 3, 4, 5

_dp = ctypes.POINTER(ctypes.c_double)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def build_oracle(ref=True):
    """Build liboracle.so (and oracle/_ref when /root/reference is present)."""
    target = "all" if ref else "oracle"
    subprocess.run(["make", "-s", "-f", os.path.join(ORACLE_DIR, "Makefile"), target], check=True, cwd=ROOT)


class _RefSettings(ctypes.Structure):
    _fields_ = [
        ("n_burnin", ctypes.c_long), ("n_keep", ctypes.c_long), ("n_leap_steps", ctypes.c_long),
        ("step_size", ctypes.c_double), ("precond", ctypes.c_void_p), ("n_fp_steps", ctypes.c_long),
        ("n_adapt_draws", ctypes.c_long), ("target_accept_rate", ctypes.c_double), ("gamma_val", ctypes.c_double),
        ("t0_val", ctypes.c_double), ("kappa_val", ctypes.c_double), ("max_tree_depth", ctypes.c_long),
        ("use_nuts_defaults", ctypes.c_int), ("vals_bound", ctypes.c_int), ("lower", ctypes.c_void_p), ("upper", ctypes.c_void_p),
        ("metric_id", ctypes.c_int),
    ]


class _OracleCfg(ctypes.Structure):
    _fields_ = [
        ("sampler", ctypes.c_int), ("target_id", ctypes.c_int), ("tdata", ctypes.c_void_p), ("d", ctypes.c_int),
        ("n_burnin", ctypes.c_long), ("n_keep", ctypes.c_long), ("n_leap_steps", ctypes.c_long),
        ("step_size", ctypes.c_double), ("precond", ctypes.c_void_p), ("chol_mode", ctypes.c_int),
        ("n_fp_steps", ctypes.c_long), ("n_adapt_draws", ctypes.c_long),
        ("target_accept_rate", ctypes.c_double), ("gamma_val", ctypes.c_double), ("t0_val", ctypes.c_double),
        ("kappa_val", ctypes.c_double), ("max_tree_depth", ctypes.c_long),
        ("rng_mode", ctypes.c_int), ("seed", ctypes.c_ulong), ("tape", ctypes.c_void_p), ("tape_len", ctypes.c_long),
        ("chain_id", ctypes.c_long), ("sum_mode", ctypes.c_int), ("mala_exact_dmvnorm", ctypes.c_int),
        ("tape_out", ctypes.c_void_p), ("tape_out_cap", ctypes.c_long),
        ("vals_bound", ctypes.c_int), ("lower", ctypes.c_void_p), ("upper", ctypes.c_void_p), ("metric_id", ctypes.c_int),
        ("dense_jacobian", ctypes.c_int),
    ]


class _RefDeSettings(ctypes.Structure):
    _fields_ = [("n_pop", ctypes.c_long), ("n_burnin", ctypes.c_long), ("n_keep", ctypes.c_long), ("jumps", ctypes.c_int),
                ("par_b", ctypes.c_double), ("par_gamma_jump", ctypes.c_double), ("init_lb", ctypes.c_void_p),
                ("init_ub", ctypes.c_void_p), ("vals_bound", ctypes.c_int), ("lower", ctypes.c_void_p), ("upper", ctypes.c_void_p)]


class _DeCfg(ctypes.Structure):
    _fields_ = [("target_id", ctypes.c_int), ("tdata", ctypes.c_void_p), ("d", ctypes.c_int), ("n_pop", ctypes.c_long),
                ("n_burnin", ctypes.c_long), ("n_keep", ctypes.c_long), ("jumps", ctypes.c_int), ("par_b", ctypes.c_double),
                ("par_gamma_jump", ctypes.c_double), ("init_lb", ctypes.c_void_p), ("init_ub", ctypes.c_void_p),
                ("vals_bound", ctypes.c_int), ("lower", ctypes.c_void_p), ("upper", ctypes.c_void_p), ("rng_mode", ctypes.c_int),
                ("seed", ctypes.c_ulong), ("tape", ctypes.c_void_p), ("tape_len", ctypes.c_long), ("sum_mode", ctypes.c_int),
                ("tape_out", ctypes.c_void_p), ("tape_out_cap", ctypes.c_long)]


class DeSettings(dict):
    """de_settings_t with the reference's defaults (include/misc/mcmc_structs.hpp:44-62)."""

    DEFAULTS = dict(n_pop=100, n_burnin=1000, n_keep=1000, jumps=False, par_b=1e-4, par_gamma_jump=2.0, initial_lb=None, initial_ub=None,
                    lower_bounds=None, upper_bounds=None)

    def __init__(self, **kw):
        super().__init__(self.DEFAULTS)
        for k in kw:
            if k not in self.DEFAULTS:
                raise KeyError(k)
        self.update(kw)


class _OracleRes(ctypes.Structure):
    _fields_ = [("n_accept", ctypes.c_long), ("tape_used", ctypes.c_long), ("final_step", ctypes.c_double),
                ("n_leapfrog", ctypes.c_long)]


class Settings(dict):
    """Sampler settings with the reference's defaults (include/misc/mcmc_structs.hpp:66-134)."""

    DEFAULTS = dict(n_burnin=1000, n_keep=1000, n_leap_steps=1, step_size=1.0, precond=None, n_fp_steps=5,
                    n_adapt_draws=1000, target_accept_rate=0.55, gamma_val=0.05, t0_val=10.0, kappa_val=0.75,
                    max_tree_depth=10, lower_bounds=None, upper_bounds=None, metric_id=0)

    def __init__(self, **kw):
        super().__init__(self.DEFAULTS)
        for k in kw:
            if k not in self.DEFAULTS:
                raise KeyError(k)
        self.update(kw)


def _bounds(st, d, keep):
    """(vals_bound, lower ptr, upper ptr): vals_bound is on when either bound vector is given (+-inf = open side)."""
    if st["lower_bounds"] is None and st["upper_bounds"] is None:
        return 0, None, None
    lo = np.full(d, -np.inf) if st["lower_bounds"] is None else np.ascontiguousarray(st["lower_bounds"], dtype=np.float64)
    hi = np.full(d, np.inf) if st["upper_bounds"] is None else np.ascontiguousarray(st["upper_bounds"], dtype=np.float64)
    keep += [lo, hi]
    return 1, _ptr(lo), _ptr(hi)


def _precond_colmajor(precond):
    if precond is None:
        return None
    p = np.asarray(precond, dtype=np.float64)
    return np.ascontiguousarray(p.T).copy()  # column-major bytes of p


class Reference:
    def __init__(self, flavour="strict"):
        path = os.path.join(ORACLE_DIR, "_ref", "libmcmc_ref_%s.so" % flavour)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        self.lib.ref_run_chain.restype = ctypes.c_int
        self.lib.ref_run_chains.restype = ctypes.c_int
        self.lib.ref_max_threads.restype = ctypes.c_int

    @staticmethod
    def available(flavour="strict"):
        return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libmcmc_ref_%s.so" % flavour))

    def _settings(self, st, keep, d):
        pc = _precond_colmajor(st["precond"])
        keep.append(pc)
        vb, lo, hi = _bounds(st, d, keep)
        return _RefSettings(st["n_burnin"], st["n_keep"], st["n_leap_steps"], st["step_size"], _ptr(pc),
                            st["n_fp_steps"], st["n_adapt_draws"], st["target_accept_rate"], st["gamma_val"],
                            st["t0_val"], st["kappa_val"], st["max_tree_depth"], 0, vb, lo, hi, st["metric_id"])

    def run_chain(self, sampler, target_id, tdata, x0, st, seed):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        d = x0.size
        tdata = np.ascontiguousarray(tdata if tdata is not None else [0.0], dtype=np.float64)
        keep = []
        rs = self._settings(st, keep, d)
        draws = np.zeros((st["n_keep"], d))
        acc = ctypes.c_long(0)
        rc = self.lib.ref_run_chain(sampler, target_id, _ptr(tdata), d, _ptr(x0), ctypes.byref(rs),
                                    ctypes.c_ulong(seed), _ptr(draws), ctypes.byref(acc))
        assert rc == 0, rc
        return draws, acc.value

    def run_chains(self, sampler, target_id, tdata, x0s, st, seed_base, n_threads=0, keep_draws=True):
        x0s = np.ascontiguousarray(x0s, dtype=np.float64)
        C, d = x0s.shape
        tdata = np.ascontiguousarray(tdata if tdata is not None else [0.0], dtype=np.float64)
        keep = []
        rs = self._settings(st, keep, d)
        draws = np.zeros((C, st["n_keep"], d)) if keep_draws else None
        acc = np.zeros(C, dtype=np.int64)
        el = ctypes.c_double(0)
        rc = self.lib.ref_run_chains(sampler, target_id, _ptr(tdata), d, ctypes.c_long(C), _ptr(x0s),
                                     ctypes.byref(rs), ctypes.c_ulong(seed_base), _ptr(draws), _ptr(acc),
                                     int(n_threads), ctypes.byref(el))
        assert rc == 0, rc
        return draws, acc, el.value

    def run_de(self, target_id, tdata, x0, st, seed):
        """mcmc::de, one population, single-threaded member loop -> (draws [n_keep][n_pop][d], n_accept)."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        d = x0.size
        tdata = np.ascontiguousarray(tdata if tdata is not None else [0.0], dtype=np.float64)
        keep = []
        vb, lo, hi = _bounds(st, d, keep)
        ilb = None if st["initial_lb"] is None else np.ascontiguousarray(st["initial_lb"], dtype=np.float64)
        iub = None if st["initial_ub"] is None else np.ascontiguousarray(st["initial_ub"], dtype=np.float64)
        rs = _RefDeSettings(st["n_pop"], st["n_burnin"], st["n_keep"], int(st["jumps"]), st["par_b"], st["par_gamma_jump"], _ptr(ilb),
                            _ptr(iub), vb, lo, hi)
        draws = np.zeros((st["n_keep"], st["n_pop"], d))
        acc = ctypes.c_long(0)
        self.lib.ref_run_de.restype = ctypes.c_int
        rc = self.lib.ref_run_de(target_id, _ptr(tdata), d, _ptr(x0), ctypes.byref(rs), ctypes.c_ulong(seed), _ptr(draws), ctypes.byref(acc))
        assert rc == 0, rc
        return draws, acc.value

    def max_threads(self):
        return self.lib.ref_max_threads()

    def rng_stream(self, seed, n_norm, n_unif):
        out = np.zeros(n_norm + n_unif)
        self.lib.ref_rng_stream(ctypes.c_ulong(seed), ctypes.c_long(n_norm), ctypes.c_long(n_unif), _ptr(out))
        return out


class Oracle:
    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle(ref=False)
        self.lib = ctypes.CDLL(path)
        self.lib.oracle_run_chain.restype = ctypes.c_int
        self.lib.oracle_target.restype = ctypes.c_double

    def run_chain(self, sampler, target_id, tdata, x0, st, seed=0, rng_mode=RNG_MT, tape=None, chain_id=0,
                  sum_mode=SUM_SEQ, chol_mode=1, mala_exact=0, record_tape=0, want_logp=False, want_margins=False, dense_jacobian=0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        d = x0.size
        tdata = np.ascontiguousarray(tdata if tdata is not None else [0.0], dtype=np.float64)
        pc = _precond_colmajor(st["precond"])
        tape_a = None if tape is None else np.ascontiguousarray(tape, dtype=np.float64)
        rec = np.zeros(record_tape) if record_tape else None
        keep = []
        vb, lo, hi = _bounds(st, d, keep)
        cfg = _OracleCfg(sampler, target_id, _ptr(tdata), d, st["n_burnin"], st["n_keep"], st["n_leap_steps"],
                         st["step_size"], _ptr(pc), chol_mode, st["n_fp_steps"], st["n_adapt_draws"],
                         st["target_accept_rate"], st["gamma_val"], st["t0_val"], st["kappa_val"],
                         st["max_tree_depth"], rng_mode, ctypes.c_ulong(seed), _ptr(tape_a),
                         0 if tape_a is None else tape_a.size, chain_id, sum_mode, mala_exact, _ptr(rec),
                         record_tape, vb, lo, hi, st["metric_id"], int(dense_jacobian))
        draws = np.zeros((st["n_keep"], d))
        logp = np.zeros(st["n_keep"]) if want_logp else None
        res = _OracleRes()
        margins = np.full(st["n_burnin"] + st["n_keep"], np.nan) if want_margins else None
        if want_margins:   # u - exp(comp) of every accept test (HMC / MALA / RWMH / RM-HMC), see oracle.cpp note_margin
            self.lib.oracle_set_margin_buffer(_ptr(margins), ctypes.c_long(margins.size))
        try:
            rc = self.lib.oracle_run_chain(ctypes.byref(cfg), _ptr(x0), _ptr(draws), _ptr(logp), ctypes.byref(res))
        finally:
            if want_margins:
                self.lib.oracle_set_margin_buffer(None, ctypes.c_long(0))
        assert rc == 0, rc
        out = dict(draws=draws, n_accept=res.n_accept, tape_used=res.tape_used, final_step=res.final_step,
                   n_leapfrog=res.n_leapfrog)
        if want_logp:
            out["logp"] = logp
        if want_margins:
            out["margins"] = margins
        if record_tape:
            out["tape"] = rec[:min(record_tape, res.tape_used)]
        return out

    def run_de(self, target_id, tdata, x0, st, seed=0, rng_mode=RNG_MT, tape=None, sum_mode=SUM_SEQ, record_tape=0):
        """The restated mcmc::de (oracle.cpp oracle_run_de): draws [n_keep][n_pop][d], n_accept, optionally the variate tape."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        d = x0.size
        tdata = np.ascontiguousarray(tdata if tdata is not None else [0.0], dtype=np.float64)
        keep = []
        vb, lo, hi = _bounds(st, d, keep)
        ilb = None if st["initial_lb"] is None else np.ascontiguousarray(st["initial_lb"], dtype=np.float64)
        iub = None if st["initial_ub"] is None else np.ascontiguousarray(st["initial_ub"], dtype=np.float64)
        tape_a = None if tape is None else np.ascontiguousarray(tape, dtype=np.float64)
        rec = np.zeros(record_tape) if record_tape else None
        cfg = _DeCfg(target_id, _ptr(tdata), d, st["n_pop"], st["n_burnin"], st["n_keep"], int(st["jumps"]), st["par_b"],
                     st["par_gamma_jump"], _ptr(ilb), _ptr(iub), vb, lo, hi, rng_mode, ctypes.c_ulong(seed), _ptr(tape_a),
                     0 if tape_a is None else tape_a.size, sum_mode, _ptr(rec), record_tape)
        draws = np.zeros((st["n_keep"], st["n_pop"], d))
        res = _OracleRes()
        self.lib.oracle_run_de.restype = ctypes.c_int
        rc = self.lib.oracle_run_de(ctypes.byref(cfg), _ptr(x0), _ptr(draws), ctypes.byref(res))
        assert rc == 0, rc
        out = dict(draws=draws, n_accept=res.n_accept, tape_used=res.tape_used)
        if record_tape:
            out["tape"] = rec[:min(record_tape, res.tape_used)]
        return out

    def target(self, target_id, tdata, x, sum_mode=SUM_SEQ, want_grad=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        tdata = np.ascontiguousarray(tdata if tdata is not None else [0.0], dtype=np.float64)
        g = np.zeros_like(x) if want_grad else None
        v = self.lib.oracle_target(target_id, _ptr(tdata), x.size, _ptr(x), _ptr(g), sum_mode)
        return v, g

    def metric(self, target_id, metric_id, tdata, x, want_deriv=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        d = x.size
        tdata = np.ascontiguousarray(tdata if tdata is not None else [0.0], dtype=np.float64)
        G = np.zeros((d, d))
        dG = np.zeros((d, d, d)) if want_deriv else None
        rc = self.lib.oracle_metric(target_id, metric_id, _ptr(tdata), d, _ptr(x), _ptr(G), _ptr(dG))
        assert rc == 0, rc
        return G.T.copy(), (None if dG is None else dG.transpose(0, 2, 1).copy())   # column-major -> [i][j]

    def rng_stream(self, rng_mode, seed, chain_id, draw, d, n_unif):
        out = np.zeros(d + n_unif)
        self.lib.oracle_rng_stream(rng_mode, ctypes.c_ulong(seed), ctypes.c_long(chain_id), ctypes.c_long(draw), d,
                                   n_unif, _ptr(out))
        return out


def c2_initial(n_chains, d, first_chain=0):
    """x0[c][j] = sin(0.37 c + 0.11 j) — SURVEY.md §8(d) C2."""
    c = np.arange(first_chain, first_chain + n_chains, dtype=np.float64)[:, None]
    j = np.arange(d, dtype=np.float64)[None, :]
    return np.sin(0.37 * c + 0.11 * j)
