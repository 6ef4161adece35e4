"""GPU parity tests for the RWMH path (src/rwmh.cpp:30-199) through the C ABI."""
import numpy as np
import pytest

import golden_util
import oracle_lib as ol
from test_gpu_hmc import _oracle_chains, _sym_pd, TOL

pytestmark = pytest.mark.gpu


def test_golden_cases(engine):
    """The committed fixtures generated from the unmodified reference (tests/golden/make_golden.py): rwmh_*."""
    names = {"iso_gauss": ol.TGT_ISO_GAUSS, "diag_gauss": ol.TGT_DIAG_GAUSS, "dense_gauss": ol.TGT_DENSE_GAUSS}
    inv = {v: k for k, v in names.items()}
    seen = 0
    for c in golden_util.load()["cases"]:
        if c["sampler"] != ol.RWMH:
            continue
        st = c["settings"]
        r = engine.rwmh(np.array([c["x0"]], dtype=np.float64), inv[c["target"]], target_data=c["tdata"], par_scale=st["step_size"],
                        cov_mat=st["precond"], n_burnin=st["n_burnin"], n_keep=st["n_keep"], rng_mode=engine.api.RNG_MT19937_TAPE,
                        seed=c["seed"], arith=engine.api.ARITH_STRICT, lower_bounds=st["lower_bounds"], upper_bounds=st["upper_bounds"])
        assert np.abs(r["draws"][0] - c["draws"]).max() <= TOL, c["name"]
        assert r["n_accept"][0] == c["n_accept"], c["name"]
        if st["lower_bounds"] is None and st["precond"] is None:
            assert np.array_equal(r["draws"][0], c["draws"]), c["name"]   # no transcendental, no reduction-order freedom
        seen += 1
    assert seen == 4


@pytest.mark.parametrize("d", [3, 16, 128, 200, 512])
def test_iso_vs_reference_and_oracle(engine, reference, oracle, d):
    C = 6
    x0 = ol.c2_initial(C, d)
    scale = 1.2 / d ** 0.5
    st = ol.Settings(n_burnin=5, n_keep=80, step_size=scale)
    od, oa, olp = _oracle_chains(oracle, ol.RWMH, ol.TGT_ISO_GAUSS, None, x0, st, 61, ol.RNG_MT, ol.SUM_WARP)
    r = engine.rwmh(x0, "iso_gauss", par_scale=scale, n_burnin=5, n_keep=80, rng_mode=engine.api.RNG_MT19937_TAPE, seed=61,
                    arith=engine.api.ARITH_STRICT, want_logp=True)
    assert np.array_equal(r["draws"], od)
    assert np.array_equal(r["n_accept"], oa)
    assert np.array_equal(r["logp"], olp)
    assert 0 < oa.max() and oa.min() < 80  # accepts and rejects both exercised
    ref, acc, _ = reference.run_chains(ol.RWMH, ol.TGT_ISO_GAUSS, None, x0, st, 61)
    assert np.abs(r["draws"] - ref).max() <= TOL
    assert np.array_equal(r["n_accept"], acc)
    rf = engine.rwmh(x0, "iso_gauss", par_scale=scale, n_burnin=5, n_keep=80, rng_mode=engine.api.RNG_MT19937_TAPE, seed=61,
                     arith=engine.api.ARITH_FAST)
    assert np.abs(rf["draws"] - od).max() <= TOL
    assert np.array_equal(rf["n_accept"], oa)


@pytest.mark.parametrize("chol_mode", [0, 1])
def test_dense_cov_dense_target(engine, reference, oracle, chol_mode):
    rng = np.random.default_rng(17)
    d, C = 24, 5
    P = _sym_pd(rng, d, 1.0)
    cov = _sym_pd(rng, d, 0.5)
    x0 = rng.normal(size=(C, d))
    st = ol.Settings(n_burnin=5, n_keep=100, step_size=0.2, precond=cov)
    od, oa, _ = _oracle_chains(oracle, ol.RWMH, ol.TGT_DENSE_GAUSS, P.ravel(), x0, st, 9, ol.RNG_MT, ol.SUM_WARP, chol_mode=chol_mode)
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = engine.rwmh(x0, "dense_gauss", target_data=P.ravel(), par_scale=0.2, cov_mat=cov, n_burnin=5, n_keep=100,
                        rng_mode=engine.api.RNG_MT19937_TAPE, seed=9, arith=arith, chol_mode=chol_mode)
        assert np.abs(r["draws"] - od).max() <= TOL
        assert np.array_equal(r["n_accept"], oa)
    if chol_mode == 1:   # the reference's Eigen backend keeps the upper triangle of cov in its "lower" factor
        ref, acc, _ = reference.run_chains(ol.RWMH, ol.TGT_DENSE_GAUSS, P.ravel(), x0, st, 9)
        assert np.abs(r["draws"] - ref).max() <= TOL
        assert np.array_equal(r["n_accept"], acc)


def test_box_constraints(engine, reference, oracle):
    """vals_bound with every bound type, identity and dense proposal covariance (src/rwmh.cpp:82-93,105-107,158-165)."""
    inf = np.inf
    d = 10
    lo = np.array([-inf, 0.0, -inf, -1.0, -2.0, -inf, 0.5, -inf, -3.0, -inf])
    hi = np.array([inf, inf, 2.0, 1.5, 2.0, inf, inf, 0.0, 3.0, 4.0])
    x0 = np.tile(np.array([0.1, 0.4, 1.0, 0.2, -0.5, 0.3, 1.5, -0.7, 0.0, 1.0]), (4, 1)) * np.linspace(0.8, 1.1, 4)[:, None]
    w = np.linspace(0.6, 1.6, d)
    for cov in (None, np.diag(np.linspace(0.5, 2.0, d)) + 0.1):
        st = ol.Settings(n_burnin=5, n_keep=80, step_size=0.3, precond=cov, lower_bounds=lo, upper_bounds=hi)
        ref, acc, _ = reference.run_chains(ol.RWMH, ol.TGT_DIAG_GAUSS, w, x0, st, 71)
        for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
            r = engine.rwmh(x0, "diag_gauss", target_data=w, par_scale=0.3, cov_mat=cov, n_burnin=5, n_keep=80,
                            rng_mode=engine.api.RNG_MT19937_TAPE, seed=71, arith=arith, lower_bounds=lo, upper_bounds=hi)
            assert np.abs(r["draws"] - ref).max() <= TOL
            assert np.array_equal(r["n_accept"], acc)
            assert np.all(r["draws"] >= lo) and np.all(r["draws"] <= hi)


def test_philox_and_sharding(engine, oracle):
    C, d = 12, 96
    x0 = ol.c2_initial(C, d)
    st = ol.Settings(n_burnin=3, n_keep=40, step_size=0.12)
    od, oa, _ = _oracle_chains(oracle, ol.RWMH, ol.TGT_ISO_GAUSS, None, x0, st, 4243, ol.RNG_PHILOX, ol.SUM_WARP)
    r = engine.rwmh(x0, "iso_gauss", par_scale=0.12, n_burnin=3, n_keep=40, rng_mode=engine.api.RNG_PHILOX, seed=4243)
    assert np.abs(r["draws"] - od).max() <= TOL and np.array_equal(r["n_accept"], oa)
    part = engine.rwmh(x0[8:], "iso_gauss", par_scale=0.12, n_burnin=3, n_keep=40, rng_mode=engine.api.RNG_PHILOX, seed=4243,
                       chain_offset=8)
    assert np.array_equal(part["draws"], r["draws"][8:])


def test_many_chains_moments_and_edge_cases(engine):
    """4096 chains at d=64: stationary moments; zero kept draws / zero burn-in are accepted like the reference."""
    rng = np.random.default_rng(2)
    d, C = 64, 4096
    w = np.linspace(0.5, 2.0, d)
    x0 = rng.normal(size=(C, d)) / np.sqrt(w)
    r = engine.rwmh(x0, "diag_gauss", target_data=w, par_scale=0.25, n_burnin=400, n_keep=20, rng_mode=engine.api.RNG_PHILOX, seed=5)
    v = r["draws"].var(axis=(0, 1))
    assert np.abs(v * w - 1).max() < 0.12
    assert 0.1 < r["n_accept"].mean() / 20 < 0.6
    r0 = engine.rwmh(x0[:3], "diag_gauss", target_data=w, par_scale=0.25, n_burnin=7, n_keep=0, rng_mode=engine.api.RNG_PHILOX, seed=5)
    assert r0["draws"].shape == (3, 0, d) and (r0["n_accept"] == 0).all()
