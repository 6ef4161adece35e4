"""GPU parity tests for the MALA path (src/mala.cpp + mala.ipp + dmvnorm.hpp) through the C ABI."""
import numpy as np
import pytest

import oracle_lib as ol
from test_gpu_hmc import _oracle_chains, _sym_pd, TOL

pytestmark = pytest.mark.gpu


def test_g3_d3_vs_reference(engine, reference):
    """d=3 standard Gaussian, seed 1, eps=0.5."""
    st = ol.Settings(n_burnin=0, n_keep=5, step_size=0.5)
    ref, acc = reference.run_chain(ol.MALA, ol.TGT_ISO_GAUSS, None, [1, -1, 0.5], st, 1)
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = engine.mala(np.array([[1, -1, 0.5]]), "iso_gauss", step_size=0.5, n_burnin=0, n_keep=5,
                        rng_mode=engine.api.RNG_MT19937_TAPE, seed=1, arith=arith)
        assert np.abs(r["draws"][0] - ref).max() <= TOL
        assert r["n_accept"][0] == acc


@pytest.mark.parametrize("d", [16, 128, 200, 512])
def test_iso_vs_reference_and_oracle(engine, reference, oracle, d):
    C = 6
    x0 = ol.c2_initial(C, d)
    eps = 0.9 / d ** 0.25
    st = ol.Settings(n_burnin=5, n_keep=60, step_size=eps)
    od, oa, olp = _oracle_chains(oracle, ol.MALA, ol.TGT_ISO_GAUSS, None, x0, st, 31, ol.RNG_MT, ol.SUM_WARP)
    r = engine.mala(x0, "iso_gauss", step_size=eps, n_burnin=5, n_keep=60, rng_mode=engine.api.RNG_MT19937_TAPE,
                    seed=31, arith=engine.api.ARITH_STRICT, want_logp=True)
    assert np.array_equal(r["draws"], od)
    assert np.array_equal(r["n_accept"], oa)
    assert np.array_equal(r["logp"], olp)
    assert 0 < oa.max() and oa.min() < 60  # accepts and rejects both exercised
    if d <= 128:  # the literal reference is O(d^3) per draw
        ref, acc, _ = reference.run_chains(ol.MALA, ol.TGT_ISO_GAUSS, None, x0, st, 31)
        assert np.abs(r["draws"] - ref).max() <= TOL
        assert np.array_equal(r["n_accept"], acc)
    rf = engine.mala(x0, "iso_gauss", step_size=eps, n_burnin=5, n_keep=60, rng_mode=engine.api.RNG_MT19937_TAPE,
                     seed=31, arith=engine.api.ARITH_FAST)
    assert np.abs(rf["draws"] - od).max() <= TOL


@pytest.mark.parametrize("chol_mode", [0, 1])
def test_linreg_dense_precond(engine, reference, oracle, chol_mode):
    """C3-shaped case at a size the CPU reference can do: Bayesian linear regression posterior, dense M."""
    rng = np.random.default_rng(7)
    d, C = 24, 5
    A = _sym_pd(rng, d, 2.0); b = rng.normal(size=d)
    M = _sym_pd(rng, d, 0.5)
    td = np.concatenate([A.ravel(), b])
    x0 = rng.normal(size=(C, d))
    st = ol.Settings(n_burnin=5, n_keep=80, step_size=0.25, precond=M)
    od, oa, _ = _oracle_chains(oracle, ol.MALA, ol.TGT_LINREG, td, x0, st, 8, ol.RNG_MT, ol.SUM_WARP, chol_mode=chol_mode)
    r = engine.mala(x0, "linreg", target_data=td, step_size=0.25, precond_mat=M, n_burnin=5, n_keep=80,
                    rng_mode=engine.api.RNG_MT19937_TAPE, seed=8, arith=engine.api.ARITH_STRICT, chol_mode=chol_mode)
    assert np.abs(r["draws"] - od).max() <= TOL
    assert np.array_equal(r["n_accept"], oa)
    if chol_mode == 1:
        ref, acc, _ = reference.run_chains(ol.MALA, ol.TGT_LINREG, td, x0, st, 8)
        assert np.abs(r["draws"] - ref).max() <= TOL
        assert np.array_equal(r["n_accept"], acc)


def test_philox_and_sharding(engine, oracle):
    C, d = 12, 96
    x0 = ol.c2_initial(C, d)
    st = ol.Settings(n_burnin=3, n_keep=30, step_size=0.3)
    od, oa, _ = _oracle_chains(oracle, ol.MALA, ol.TGT_ISO_GAUSS, None, x0, st, 4242, ol.RNG_PHILOX, ol.SUM_WARP)
    r = engine.mala(x0, "iso_gauss", step_size=0.3, n_burnin=3, n_keep=30, rng_mode=engine.api.RNG_PHILOX, seed=4242)
    assert np.abs(r["draws"] - od).max() <= TOL and np.array_equal(r["n_accept"], oa)
    part = engine.mala(x0[8:], "iso_gauss", step_size=0.3, n_burnin=3, n_keep=30, rng_mode=engine.api.RNG_PHILOX, seed=4242,
                       chain_offset=8)
    assert np.array_equal(part["draws"], r["draws"][8:])


def test_many_chains_moments(engine):
    """2048 chains of the C3-family target at d=256: stationary moments of the Gaussian posterior."""
    rng = np.random.default_rng(1)
    d, C = 256, 2048
    w = np.linspace(0.5, 2.0, d)
    x0 = rng.normal(size=(C, d)) / np.sqrt(w)
    r = engine.mala(x0, "diag_gauss", target_data=w, step_size=0.25, n_burnin=200, n_keep=50, rng_mode=engine.api.RNG_PHILOX,
                    seed=99)
    assert np.isfinite(r["draws"]).all()
    v = r["draws"].var(axis=(0, 1))
    assert np.abs(v * w - 1).max() < 0.12  # 2048 chains: sd of a variance estimate ~3%, max over 256 coordinates
    assert 0.2 < r["n_accept"].mean() / 50 < 0.99
