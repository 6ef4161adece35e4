"""C-ABI behaviour on a GPU box: error codes, device functors vs host callbacks, library-generated tape."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def test_error_codes(engine):
    api = engine.api
    with pytest.raises(engine.McmcB200Error) as e:
        engine.hmc(np.zeros((2, 8)), 99, n_burnin=1, n_keep=1)
    assert e.value.code == api.ERR_UNKNOWN_TARGET
    with pytest.raises(engine.McmcB200Error) as e:
        engine.hmc(np.zeros((2, 8)), "diag_gauss", n_burnin=1, n_keep=1)  # data blob missing
    assert e.value.code == api.ERR_INVALID_ARG
    with pytest.raises(engine.McmcB200Error) as e:
        engine.hmc(np.zeros((2, 2100)), "iso_gauss", n_burnin=1, n_keep=1)  # beyond every HMC kernel (max 2048)
    assert e.value.code == api.ERR_UNSUPPORTED
    with pytest.raises(engine.McmcB200Error) as e:
        engine.hmc(np.zeros((2, 4)), "iso_gauss", n_burnin=1, n_keep=1, precond_mat=-np.eye(4))
    assert e.value.code == api.ERR_INVALID_ARG
    with pytest.raises(engine.McmcB200Error) as e:
        engine.hmc(np.zeros((2, 4)), "iso_gauss", n_burnin=1, n_keep=1, device=63)
    assert e.value.code == api.ERR_INVALID_ARG
    # zero kept draws is legal (reference: draws_out resized to 0 x d)
    r = engine.hmc(np.zeros((2, 4)), "iso_gauss", n_burnin=3, n_keep=0)
    assert r["draws"].shape == (2, 0, 4)


@pytest.mark.parametrize("tname,tid", [("iso_gauss", 0), ("diag_gauss", 1), ("dense_gauss", 2), ("linreg", 3), ("normal_model", 4),
                                       ("funnel", 5)])
def test_device_functors_match_host_callbacks(engine, oracle, tname, tid):
    rng = np.random.default_rng(tid)
    d = 2 if tid == 4 else 37
    if tid == 1:
        td = rng.uniform(0.5, 2, size=d)
    elif tid == 2:
        a = rng.normal(size=(d, d)); P = a @ a.T + np.eye(d); td = ((P + P.T) / 2).ravel()
    elif tid == 3:
        a = rng.normal(size=(d, d)); P = a @ a.T + np.eye(d); td = np.concatenate([((P + P.T) / 2).ravel(), rng.normal(size=d)])
    elif tid == 4:
        td = np.array([100.0, 2.0, 197.0])
    else:
        td = None
    x = rng.normal(size=(9, d)) + (2.5 if tid == 4 else 0.0)
    val, grad = engine.api.target_eval(tname, td, x, arith=engine.api.ARITH_STRICT)
    for i in range(x.shape[0]):
        v, g = oracle.target(tid, td, x[i], sum_mode=ol.SUM_WARP)
        if tid in (4, 5):  # device log() / exp() vs glibc: last-bit differences allowed
            assert abs(val[i] - v) <= 1e-12 * abs(v) and np.abs(grad[i] - g).max() <= 1e-12 * np.abs(g).max()
        else:
            assert val[i] == v and np.array_equal(grad[i], g)


def test_library_tape_matches_oracle_stream(engine, oracle):
    """mcmcb200_mt19937_tape (host side of MT19937 mode) vs the variates the oracle consumes from std::mt19937_64."""
    st = ol.Settings(n_burnin=2, n_keep=3, n_leap_steps=1, step_size=0.1)
    o = oracle.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, np.zeros(5), st, seed=42, record_tape=100)
    assert np.array_equal(engine.api.mt19937_tape(42, 0, 5, 5), o["tape"])
    xs = 2 + 2 * np.sin(np.arange(100.0))
    td = np.array([100.0, xs.mean(), ((xs - xs.mean()) ** 2).sum()])
    o = oracle.run_chain(ol.RMHMC, ol.TGT_NORMAL_MODEL, td, [3, 3], st, seed=43, record_tape=100)
    assert np.array_equal(engine.api.mt19937_tape(43, 2, 5, 2), o["tape"])


def test_run_calls_are_reentrant_across_host_threads(engine):
    """Like the reference's samplers (no globals), the run calls may be issued from several host threads at once, on
    the same device: every thread has its own device scratch.  Concurrent results equal the sequential ones."""
    import threading

    d, C = 128, 64
    x0 = ol.c2_initial(4 * C, d)
    kw = dict(n_leap_steps=10, step_size=0.1, n_burnin=20, n_keep=50, rng_mode=engine.api.RNG_PHILOX, seed=9)
    want = [engine.hmc(x0[i * C:(i + 1) * C], "iso_gauss", chain_offset=i * C, **kw)["draws"].copy() for i in range(4)]
    got = [None] * 4

    def work(i):
        for _ in range(5):
            got[i] = engine.hmc(x0[i * C:(i + 1) * C], "iso_gauss", chain_offset=i * C, **kw)["draws"]

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(4):
        assert np.array_equal(got[i], want[i]), i


def test_strict_mala_never_takes_the_gemm_path_and_tapes_are_validated(engine, oracle):
    """(a) MCMCB200_ARITH_STRICT with >= 256 chains of a dense quadratic target: the chain-batched tensor-core GEMM path has no
    un-contracted operation order, so STRICT must run the warp kernel and stay bit-identical to the oracle; n_dim beyond
    the warp kernel is refused in STRICT instead of silently computed in FAST arithmetic.
    (b) USER_TAPE strides shorter than what a static-count sampler consumes are rejected up front."""
    rng = np.random.default_rng(17)
    d, C = 24, 300
    a = rng.normal(size=(d, d))
    A = a @ a.T / d + np.eye(d)
    td = np.concatenate([((A + A.T) / 2).ravel(), rng.normal(size=d)])
    x0 = rng.normal(size=(C, d)) * 0.3
    st = ol.Settings(n_burnin=2, n_keep=12, step_size=0.25)
    r = engine.mala(x0, "linreg", target_data=td, step_size=0.25, n_burnin=2, n_keep=12, rng_mode=engine.api.RNG_MT19937_TAPE, seed=70,
                    arith=engine.api.ARITH_STRICT)
    assert r["kernel_launches"] == 1   # the GEMM path issues two launches per draw
    for c in (0, 128, 255, 299):
        o = oracle.run_chain(ol.MALA, ol.TGT_LINREG, td, x0[c], st, seed=70 + c, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_WARP)
        assert np.array_equal(r["draws"][c], o["draws"]) and r["n_accept"][c] == o["n_accept"], c
    rf = engine.mala(x0, "linreg", target_data=td, step_size=0.25, n_burnin=2, n_keep=12, rng_mode=engine.api.RNG_MT19937_TAPE, seed=70)
    assert rf["kernel_launches"] > 1 and np.abs(rf["draws"] - r["draws"]).max() <= 1e-10
    d2 = 600
    with pytest.raises(engine.McmcB200Error) as e:
        engine.mala(np.zeros((4, d2)), "dense_gauss", target_data=np.eye(d2).ravel(), n_burnin=1, n_keep=1, arith=engine.api.ARITH_STRICT)
    assert e.value.code == engine.api.ERR_UNSUPPORTED
    n_total = 6
    for call in (lambda tp: engine.hmc(x0[:3], "linreg", target_data=td, n_leap_steps=2, step_size=0.1, n_burnin=1, n_keep=5, rng_mode=engine.api.RNG_USER_TAPE, tape=tp),
                 lambda tp: engine.mala(x0[:3], "linreg", target_data=td, step_size=0.1, n_burnin=1, n_keep=5, rng_mode=engine.api.RNG_USER_TAPE, tape=tp),
                 lambda tp: engine.rwmh(x0[:3], "linreg", target_data=td, par_scale=0.1, n_burnin=1, n_keep=5, rng_mode=engine.api.RNG_USER_TAPE, tape=tp)):
        call(np.full((3, n_total * (d + 1)), 0.25))   # exactly enough
        with pytest.raises(engine.McmcB200Error) as e:
            call(np.full((3, n_total * (d + 1) - 1), 0.25))
        assert e.value.code == engine.api.ERR_INVALID_ARG


def test_column_major_layout_is_the_transpose_of_the_default(engine):
    """mcmcb200_output_t::draws_layout = MCMCB200_LAYOUT_COLMAJOR hands back every chain's n_keep x n_dim matrix in the reference's
    column-major Mat_t order, transposed on the device — ragged sizes included."""
    for C, d, nk in ((5, 3, 7), (9, 128, 33), (3, 70, 100), (2, 1, 1)):
        x0 = ol.c2_initial(C, d)
        kw = dict(n_leap_steps=3, step_size=0.2, n_burnin=2, n_keep=nk, rng_mode=engine.api.RNG_PHILOX, seed=4)
        a = engine.hmc(x0, "iso_gauss", **kw)["draws"]
        b = engine.hmc(x0, "iso_gauss", layout=engine.api.LAYOUT_COLMAJOR, **kw)["draws"]
        assert np.array_equal(b.reshape(C, d, nk), a.transpose(0, 2, 1))


def test_user_defined_target_runs_every_sampler_and_matches_the_reference(engine, oracle, reference):
    """examples/user_target/normal_raw.cu: the log-likelihood of /root/reference/examples/eigen/hmc_normal.cpp:44-76 written by a
    USER as a __device__ functor on the raw observations, compiled into the user's own library and registered at load time —
    targets.cuh and libmcmc_b200.so untouched.  The reference runs the same likelihood (its sufficient-statistics form is the
    same function up to rounding), so the draws must agree to the contract tolerance; the user's Fisher metric drives RM-HMC."""
    import os
    from mcmc_b200 import build

    engine.api.load_user_library(build.USER_EXAMPLE_LIB)
    xs = 2 + 2 * np.sin(np.arange(100.0))
    raw = np.concatenate([[100.0], xs])
    suff = np.array([100.0, xs.mean(), ((xs - xs.mean()) ** 2).sum()])
    val, grad = engine.api.target_eval("normal_raw", raw, np.array([[2.5, 1.7], [1.0, 3.0]]), arith=engine.api.ARITH_STRICT)
    for i, p in enumerate(([2.5, 1.7], [1.0, 3.0])):
        v, g = oracle.target(ol.TGT_NORMAL_MODEL, suff, np.array(p))
        assert abs(val[i] - v) <= 1e-11 * abs(v) and np.abs(grad[i] - g).max() <= 1e-10 * np.abs(g).max()
    x0 = np.array([[3.0, 3.0], [2.5, 2.2], [1.5, 3.5]])
    C = x0.shape[0]
    common = dict(target_data=raw, n_burnin=5, n_keep=40, rng_mode=engine.api.RNG_MT19937_TAPE, seed=11, arith=engine.api.ARITH_STRICT)
    for smp, call, st in (
            (ol.HMC, lambda: engine.hmc(x0, "normal_raw", n_leap_steps=5, step_size=0.05, **common), ol.Settings(n_burnin=5, n_keep=40, n_leap_steps=5, step_size=0.05)),
            (ol.MALA, lambda: engine.mala(x0, "normal_raw", step_size=0.1, **common), ol.Settings(n_burnin=5, n_keep=40, step_size=0.1)),
            (ol.RWMH, lambda: engine.rwmh(x0, "normal_raw", par_scale=0.2, **common), ol.Settings(n_burnin=5, n_keep=40, step_size=0.2)),
            (ol.RMHMC, lambda: engine.rmhmc(x0, "normal_raw", n_leap_steps=2, step_size=0.15, **common), ol.Settings(n_burnin=5, n_keep=40, n_leap_steps=2, step_size=0.15))):
        r = call()
        ref, acc, _ = reference.run_chains(smp, ol.TGT_NORMAL_MODEL, suff, x0, st, 11)
        assert np.abs(r["draws"] - ref).max() <= 1e-9, (smp, np.abs(r["draws"] - ref).max())
        assert np.array_equal(r["n_accept"], acc), smp
    # NUTS on Philox: stationarity only (the oracle's Philox stream on the sufficient-statistics form differs by rounding)
    r = engine.nuts(np.tile([2.5, 2.5], (64, 1)), "normal_raw", target_data=raw, n_burnin=150, n_keep=100, n_adapt_draws=150, rng_mode=engine.api.RNG_PHILOX, seed=3)
    m = r["draws"].mean(axis=(0, 1))
    assert abs(m[0] - xs.mean()) < 0.1 and abs(m[1] - xs.std()) < 0.15, m
    # DE was not instantiated for this target (MCMCB200_USER_NO_DE): a loud error, not a fallback
    with pytest.raises(engine.McmcB200Error) as e:
        engine.de(np.array([[2.5, 2.5]]), "normal_raw", target_data=raw, n_pop=8, n_burnin=1, n_keep=1)
    assert e.value.code == engine.api.ERR_UNKNOWN_TARGET
