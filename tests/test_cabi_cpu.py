"""CPU tests of the product's C-ABI library (no GPU, no compute calls): it loads, exports every symbol that
include/mcmc_b200.h declares, reports the reference's defaults, keeps its registry, generates the reference random
stream on the host, and FAILS LOUDLY (no fallback) when no CUDA device is usable."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    from mcmc_b200 import api as a

    a.load()
    return a


def test_every_declared_symbol_is_exported(api):
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))) if h.endswith(".h"))
    declared = set(re.findall(r"\b(mcmcb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 16
    lib = ctypes.CDLL(api.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(api.EXPORTED_SYMBOLS)


def test_defaults_match_reference_structs(api):
    """include/misc/mcmc_structs.hpp:66-134."""
    lib = api.load()
    h = api.HmcSettings(); lib.mcmcb200_hmc_settings_default(ctypes.byref(h))
    assert (h.n_burnin_draws, h.n_keep_draws, h.n_leap_steps, h.step_size) == (1000, 1000, 1, 1.0) and not h.precond_mat
    m = api.MalaSettings(); lib.mcmcb200_mala_settings_default(ctypes.byref(m))
    assert (m.n_burnin_draws, m.n_keep_draws, m.step_size) == (1000, 1000, 1.0)
    n = api.NutsSettings(); lib.mcmcb200_nuts_settings_default(ctypes.byref(n))
    assert (n.n_burnin_draws, n.n_keep_draws, n.n_adapt_draws, n.target_accept_rate, n.max_tree_depth, n.step_size, n.gamma_val,
            n.t0_val, n.kappa_val) == (1000, 1000, 1000, 0.55, 10, 1.0, 0.05, 10.0, 0.75)
    r = api.RmhmcSettings(); lib.mcmcb200_rmhmc_settings_default(ctypes.byref(r))
    assert (r.n_burnin_draws, r.n_keep_draws, r.n_leap_steps, r.step_size, r.n_fp_steps) == (1000, 1000, 1, 1.0, 5)
    w = api.RwmhSettings(); lib.mcmcb200_rwmh_settings_default(ctypes.byref(w))   # mcmc_structs.hpp:138-149
    assert (w.n_burnin_draws, w.n_keep_draws, w.par_scale) == (1000, 1000, 1.0) and not w.cov_mat


def test_target_registry(api):
    lib = api.load()
    names = ["iso_gauss", "diag_gauss", "dense_gauss", "linreg", "normal_model", "funnel"]
    assert [lib.mcmcb200_target_lookup(n.encode()) for n in names] == [0, 1, 2, 3, 4, 5]
    assert lib.mcmcb200_target_lookup(b"nope") == -1
    f = lib.mcmcb200_target_data_len
    assert [f(t, 16) for t in range(4)] == [0, 16, 256, 272]
    assert f(4, 2) == 3 and f(4, 3) == -1 and f(9, 4) == -1
    assert f(5, 64) == 0 and f(5, 1) == -1


def test_host_side_reference_stream(api, oracle):
    """MT19937 mode's host half: the tape equals what the oracle (bit-equal to the reference) consumes."""
    st = ol.Settings(n_burnin=3, n_keep=4, n_leap_steps=1, step_size=0.1)
    o = oracle.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, np.zeros(7), st, seed=2024, record_tape=200)
    assert np.array_equal(api.mt19937_tape(2024, 0, 7, 7), o["tape"])
    o = oracle.run_chain(ol.MALA, ol.TGT_ISO_GAUSS, None, np.zeros(3), st, seed=5, record_tape=200)
    assert np.array_equal(api.mt19937_tape(5, 0, 7, 3), o["tape"])
    o = oracle.run_chain(ol.RWMH, ol.TGT_ISO_GAUSS, None, np.zeros(5), st, seed=6, record_tape=200)
    assert np.array_equal(api.mt19937_tape(6, 0, 7, 5), o["tape"])


def test_no_cpu_fallback(api):
    """Without a usable CUDA device every run call returns MCMCB200_ERR_CUDA — it never computes on the host."""
    import mcmc_b200

    if api.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    for fn in (mcmc_b200.hmc, mcmc_b200.mala, mcmc_b200.nuts, mcmc_b200.rwmh):
        with pytest.raises(mcmc_b200.McmcB200Error) as e:
            fn(np.zeros((2, 4)), "iso_gauss", n_burnin=1, n_keep=1)
        assert e.value.code == api.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or execute anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mcmc_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("oracle/oracle.cpp", "").replace(
                    "oracle/host_targets.hpp", ""), os.path.join(dirpath, f)


def test_de_reference_stream_matches_the_oracle_tape(oracle):
    """Host side of MCMCB200_RNG_MT19937_TAPE for mcmc::de (host_tape.cpp host_de_tape, no GPU needed): exactly the variates
    the restated sampler — bit-identical to the unmodified src/de.cpp — consumes, in order."""
    import numpy as np
    import oracle_lib as ol
    from mcmc_b200 import api
    from test_oracle_vs_reference import _de_cases

    for name, tid, tdata, x0, st, seed in _de_cases():
        o = oracle.run_de(tid, tdata, x0, st, seed=seed, record_tape=2_000_000)
        t = api.de_tape(seed, st["n_pop"], len(x0), st["n_burnin"] + st["n_keep"], st["par_b"])
        assert t.size == o["tape_used"] and np.array_equal(t, o["tape"]), name


def test_user_target_library_registers_without_a_gpu():
    """examples/user_target/normal_raw.cu (a USER-defined functor + metric, built by mcmc_b200.build into its own shared
    library) registers itself with libmcmc_b200.so at load time; the sampler kernels themselves need a GPU."""
    import os
    from mcmc_b200 import api, build

    assert os.path.exists(build.USER_EXAMPLE_LIB), "run __graft_entry__.build()"
    api.load_user_library(build.USER_EXAMPLE_LIB)
    tid = api.target_id("normal_raw")
    assert tid >= 64
    assert api.load().mcmcb200_target_data_len(tid, 2) == 1 and api.load().mcmcb200_target_data_len(tid, 3) == -1
    assert api.metric_lookup("funnel_softabs") == (api.TARGET_FUNNEL, 2)


def test_argument_validation_precedes_the_device(api):
    """Every run call validates its arguments before it touches the device, so a malformed call gets the specific error
    (code + text in mcmcb200_last_error) with or without a GPU — never a CUDA error, never a silent clamp.  Mirrors the
    reference's habit of refusing before sampling (src/de.cpp:83-91 bounds/size checks; include/misc/mcmc_structs.hpp
    defaults)."""
    import mcmc_b200

    x0 = np.zeros((2, 4))

    def err(fn, *a, **k):
        with pytest.raises(mcmc_b200.McmcB200Error) as e:
            fn(*a, **k)
        return e.value.code, str(e.value)

    # unknown target id / a target whose data blob is too short / n_dim invalid for the target
    code, msg = err(mcmc_b200.hmc, x0, 4242, n_burnin=1, n_keep=1)
    assert code == api.ERR_UNKNOWN_TARGET and "4242" in msg
    code, msg = err(mcmc_b200.hmc, x0, "diag_gauss", target_data=np.ones(3), n_burnin=1, n_keep=1)
    assert code == api.ERR_INVALID_ARG and "needs 4 doubles" in msg
    code, msg = err(mcmc_b200.hmc, np.zeros((2, 3)), "normal_model", target_data=np.ones(3), n_burnin=1, n_keep=1)
    assert code == api.ERR_UNKNOWN_TARGET   # the normal model has exactly two parameters
    # draw counts and trajectory lengths out of range
    code, msg = err(mcmc_b200.hmc, x0, "iso_gauss", n_burnin=-1, n_keep=1)
    assert code == api.ERR_INVALID_ARG and "draw counts" in msg
    code, msg = err(mcmc_b200.hmc, x0, "iso_gauss", n_leap_steps=-3, n_burnin=1, n_keep=1)
    assert code == api.ERR_INVALID_ARG and "n_leap_steps" in msg
    code, msg = err(mcmc_b200.rmhmc, np.zeros((2, 2)), "normal_model", target_data=np.ones(3), n_leap_steps=2**31, n_burnin=1, n_keep=1)
    assert code == api.ERR_INVALID_ARG and "n_leap_steps" in msg
    code, msg = err(mcmc_b200.nuts, x0, "iso_gauss", max_tree_depth=21, n_burnin=1, n_keep=1)
    assert code == api.ERR_INVALID_ARG and "max_tree_depth" in msg
    # n_dim beyond what the sampler/target combination supports is refused, not truncated
    code, msg = err(mcmc_b200.hmc, np.zeros((2, 4096)), "iso_gauss", n_burnin=1, n_keep=1)
    assert code == api.ERR_UNSUPPORTED and "n_dim=4096" in msg
    # mcmc::de needs two other members for a proposal (src/de.cpp:166-178) and has no log-density output
    code, msg = err(mcmc_b200.de, x0, "iso_gauss", n_pop=2, n_burnin=1, n_keep=1)
    assert code == api.ERR_INVALID_ARG and "n_pop" in msg
    code, msg = err(mcmc_b200.de, x0, "iso_gauss", n_pop=8, n_burnin=1, n_keep=1, want_logp=True)
    assert code == api.ERR_UNSUPPORTED and "logp_out" in msg
    # null pointers through the raw C ABI
    lib = api.load()
    assert lib.mcmcb200_hmc_run(None, None, None, None) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_nuts_run(None, None, None, None) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_de_run(None, None, None, None) == api.ERR_INVALID_ARG


def test_integration_note_names_every_entry_point(api):
    """INTEGRATION.md §1 maps every exported C entry point to the reference interface it replaces (or marks it new)."""
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for sym in api.EXPORTED_SYMBOLS:
        stem = sym
        if sym.endswith("_settings_default"):
            stem = "mcmcb200_*_settings_default"
        elif sym.startswith("mcmcb200_comm_"):
            stem = "mcmcb200_comm_*"
        elif sym == "mcmcb200_host_free":
            stem = "mcmcb200_host_alloc / _free"
        assert stem in txt, sym


def test_hot_kernels_keep_their_register_budget():
    """Static guard (cuobjdump on the objects build() leaves behind, no GPU): the kernels whose occupancy the measured numbers
    rest on keep their register budget and do not spill — the headline HMC kernel 72 registers / 0 stack (7 warps per
    sub-partition), the persistent NUTS kernel with 16 chains per CTA 128 registers (§4.13), the RM-HMC CTA kernel
    128 registers (4 CTAs per SM, §4.4), two-chains-per-warp HMC 64."""
    import subprocess
    import sys

    build = os.path.join(ROOT, "mcmc_b200", "build")
    if not os.path.isdir(build) or not os.path.exists(os.path.join(build, "hmc.cu.t0.o")):
        pytest.skip("no build objects here (the library was built elsewhere)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "resource_usage.py")], capture_output=True, text=True, check=True).stdout
    rows = {}
    for line in out.splitlines()[1:]:
        f = line.split(None, 5)
        rows[f[5].strip()] = tuple(int(v) for v in f[1:5])   # regs, stack, shared, local

    regs, stack, _, local = rows["hmc_pipe_kernel<IsoGauss, 4, 10, 1, false>"]
    assert regs <= 72 and stack == 0 and local == 0
    for k, v in rows.items():
        if k.startswith("nuts_pc_kernel<8,") and k.endswith(", 16>"):
            assert v[0] <= 128 and v[3] == 0, (k, v)
        if k.startswith("rmhmc_cta_kernel<Funnel, FunnelSoftabsCta") and k.endswith("false>"):
            assert v[0] <= 128 and v[3] == 0, (k, v)
        if k.startswith("hmc_half_kernel<"):
            assert v[0] <= 64 and v[1] == 0, (k, v)


def test_host_side_preconditioner_algebra(api):
    """csrc/host_linalg.cpp (no GPU): the one-time inverse and Cholesky factor of precond_mat / cov_mat that every dense-mass run
    uploads (BMO_MATOPS_INV / BMO_MATOPS_CHOL_LOWER, src/hmc.cpp:58-59).  Against numpy; chol_mode 1 keeps precond_mat's strict
    upper triangle like Eigen's matrixLLT(), chol_mode 0 zeroes it; not-positive-definite / singular input is reported,
    not factorised."""
    lib = ctypes.CDLL(api.LIB_PATH)
    inv_fn = getattr(lib, "_ZN8mcmcb20021host_inverse_colmajorEPKdiPd")
    chol_fn = getattr(lib, "_ZN8mcmcb20022host_cholesky_colmajorEPKdiiPd")
    inv_fn.restype = chol_fn.restype = ctypes.c_bool
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 33, 128):
        a = rng.normal(size=(n, n))
        M = a @ a.T / n + np.eye(n)
        M = (M + M.T) / 2
        Mc = np.asfortranarray(M)                        # column-major bytes, as the C ABI takes precond_mat
        out = np.zeros((n, n), order="F")
        assert inv_fn(Mc.ctypes.data_as(dp), n, out.ctypes.data_as(dp))
        assert np.abs(out @ M - np.eye(n)).max() <= 1e-12 * np.linalg.cond(M)
        for mode in (0, 1):
            L = np.zeros((n, n), order="F")
            assert chol_fn(Mc.ctypes.data_as(dp), n, mode, L.ctypes.data_as(dp))
            low = np.tril(L)
            assert np.abs(low @ low.T - M).max() <= 1e-13 * np.abs(M).max() * n
            assert np.abs(low - np.linalg.cholesky(M)).max() <= 1e-12
            upper = L[np.triu_indices(n, 1)]
            assert np.array_equal(upper, M[np.triu_indices(n, 1)] if mode == 1 else np.zeros_like(upper))
    # a general (non-symmetric) matrix needs the row exchanges
    G = np.asfortranarray(np.array([[0.0, 2.0, 1.0], [1.0, 0.0, 3.0], [4.0, 1.0, 0.0]]))
    out = np.zeros((3, 3), order="F")
    assert inv_fn(G.ctypes.data_as(dp), 3, out.ctypes.data_as(dp)) and np.abs(out @ G - np.eye(3)).max() <= 1e-14
    bad = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 1.0]]))           # indefinite
    sing = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 4.0]]))          # singular
    tmp = np.zeros((2, 2), order="F")
    assert not chol_fn(bad.ctypes.data_as(dp), 2, 1, tmp.ctypes.data_as(dp))
    assert not inv_fn(sing.ctypes.data_as(dp), 2, tmp.ctypes.data_as(dp))


def test_null_and_degenerate_arguments_never_crash(api):
    """Every entry point that can be reached without a device answers null pointers and degenerate sizes with an error code
    (and a message), never a crash — what a binding in another language relies on."""
    c = ctypes
    lib = api.load()
    lib.mcmcb200_last_error.restype = c.c_char_p
    lib.mcmcb200_target_data_len.restype = c.c_int64
    buf = np.zeros(16)
    p = buf.ctypes.data_as(c.c_void_p)
    assert lib.mcmcb200_target_lookup(None) == -1 and lib.mcmcb200_target_lookup(b"") == -1
    assert [lib.mcmcb200_target_data_len(2, -5), lib.mcmcb200_target_data_len(-1, 5), lib.mcmcb200_target_data_len(2, 0)] == [-1, -1, -1]
    assert lib.mcmcb200_metric_lookup(None, None, None) != 0 and b"unknown metric" in lib.mcmcb200_last_error()
    assert lib.mcmcb200_register_target(None, None) < 0
    assert lib.mcmcb200_mt19937_tape(c.c_uint64(1), c.c_int64(0), c.c_int64(0), 4, None) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_mt19937_tape(c.c_uint64(1), c.c_int64(0), c.c_int64(0), 4, p) == 0          # nothing to generate is not an error
    assert lib.mcmcb200_mt19937_tape(c.c_uint64(1), c.c_int64(-1), c.c_int64(1), 4, p) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_de_tape(c.c_uint64(1), c.c_int64(2), 4, c.c_int64(1), c.c_double(1e-4), p) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_philox_stream(c.c_uint64(1), c.c_int64(0), c.c_int64(0), 4, 3, None) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_target_eval(0, None, c.c_int64(0), 4, None, c.c_int64(0), None, None, 0) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_comm_unique_id(None, c.c_size_t(0)) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_comm_init(None, c.c_size_t(0), 2, 0, 0, None) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_comm_destroy(None) == 0
    assert lib.mcmcb200_allgather_draws(None, None, None, c.c_int64(1), 4, None, None) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_fp64_peak(0, None) == api.ERR_INVALID_ARG
    assert lib.mcmcb200_summarize_draws(None, 0, c.c_int64(1), c.c_int64(1), 4, -1, None, None) == api.ERR_INVALID_ARG
    lib.mcmcb200_host_free(None)
    lib.mcmcb200_release_workspace()
