"""GPU parity tests for box constraints (algo_settings_t::vals_bound / lower_bounds / upper_bounds) through the C ABI.

The reference runs a bounded chain in the transformed space (include/misc/transform_vals.hpp), adds the log-Jacobian to
log pi, scales the force by the diagonal "inverse Jacobian"  and maps the stored draws back.  The kernels do
the same with the CUDA math library's exp/log, so bounded parity is held to the contract tolerance 1e-10 (not bit for
bit) with identical accept counts, against (a) the committed golden vectors of the unmodified reference and (b) the
oracle on seeded many-chain configurations covering every bound type and every register-tile width."""
import numpy as np
import pytest

import golden_util
import oracle_lib as ol
from test_gpu_hmc import _oracle_chains, TOL
from test_gpu_nuts import _run_pair

pytestmark = pytest.mark.gpu

TNAME = {ol.TGT_FUNNEL: "funnel", ol.TGT_ISO_GAUSS: "iso_gauss", ol.TGT_DIAG_GAUSS: "diag_gauss", ol.TGT_DENSE_GAUSS: "dense_gauss", ol.TGT_LINREG: "linreg",
         ol.TGT_NORMAL_MODEL: "normal_model"}


def _mixed_bounds(d, rng):
    """Every bound type (none / lower / upper / both) in one vector, and a start strictly inside the box."""
    kind = np.arange(d) % 4
    lo = np.where((kind == 1) | (kind == 3), -rng.uniform(0.5, 2.0, d), -np.inf)
    hi = np.where((kind == 2) | (kind == 3), rng.uniform(0.5, 2.0, d), np.inf)
    return lo, hi


def _start(C, d, lo, hi, rng):
    x0 = rng.uniform(-0.4, 0.4, size=(C, d))
    assert np.all(x0 > lo) and np.all(x0 < hi)
    return x0


def _engine_run(engine, sampler, tid, tdata, x0s, st, arith, **kw):
    common = dict(target_data=tdata, n_burnin=st["n_burnin"], n_keep=st["n_keep"], arith=arith, want_logp=True,
                  lower_bounds=st["lower_bounds"], upper_bounds=st["upper_bounds"], **kw)
    if sampler == ol.HMC:
        return engine.hmc(x0s, TNAME[tid], n_leap_steps=st["n_leap_steps"], step_size=st["step_size"], **common)
    if sampler == ol.MALA:
        return engine.mala(x0s, TNAME[tid], step_size=st["step_size"], **common)
    if sampler == ol.RMHMC:
        return engine.rmhmc(x0s, TNAME[tid], n_leap_steps=st["n_leap_steps"], step_size=st["step_size"], n_fp_steps=st["n_fp_steps"], **common)
    if sampler == ol.RWMH:
        return engine.rwmh(x0s, TNAME[tid], par_scale=st["step_size"], cov_mat=st["precond"], **common)
    raise ValueError(sampler)


def test_bounded_golden_vectors_of_the_reference(engine, oracle):
    g = golden_util.load()
    cases = [c for c in g["cases"] if "lower" in c]
    assert {c["name"] for c in cases} == {"hmc_box_d4", "mala_box_d4", "nuts_box_d4", "rmhmc_box_sigma_positive", "rwmh_box_d4"}
    for c in cases:
        st = c["settings"]
        x0 = np.array([c["x0"]], dtype=np.float64)
        for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
            if c["sampler"] == ol.NUTS:
                r, _, _ = _run_pair(engine, oracle, c["target"], TNAME[c["target"]], c["tdata"], x0, st, c["seed"], arith)
            else:
                r = _engine_run(engine, c["sampler"], c["target"], c["tdata"], x0, st, arith, rng_mode=engine.api.RNG_MT19937_TAPE,
                                seed=c["seed"])
            assert np.abs(r["draws"][0] - c["draws"]).max() <= TOL, (c["name"], np.abs(r["draws"][0] - c["draws"]).max())
            assert r["n_accept"][0] == c["n_accept"], c["name"]
            assert np.all(r["draws"][0] >= st["lower_bounds"]) and np.all(r["draws"][0] <= st["upper_bounds"])


@pytest.mark.parametrize("d", [6, 64, 128, 200, 300, 512])
def test_bounded_hmc_many_chains_vs_oracle(engine, oracle, d):
    rng = np.random.default_rng(100 + d)
    C = 12
    lo, hi = _mixed_bounds(d, rng)
    x0 = _start(C, d, lo, hi, rng)
    w = np.exp(rng.uniform(-0.5, 0.5, d))
    st = ol.Settings(n_burnin=3, n_keep=25, n_leap_steps=5, step_size=0.05, lower_bounds=lo, upper_bounds=hi)
    od, oa, olp = _oracle_chains(oracle, ol.HMC, ol.TGT_DIAG_GAUSS, w, x0, st, 777, ol.RNG_MT, ol.SUM_WARP)
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = _engine_run(engine, ol.HMC, ol.TGT_DIAG_GAUSS, w, x0, st, arith, rng_mode=engine.api.RNG_MT19937_TAPE, seed=777)
        assert np.abs(r["draws"] - od).max() <= TOL, np.abs(r["draws"] - od).max()
        assert np.abs(r["logp"] - olp).max() <= 1e-9 * max(1.0, np.abs(olp).max())
        assert np.array_equal(r["n_accept"], oa)
    op, oap, _ = _oracle_chains(oracle, ol.HMC, ol.TGT_DIAG_GAUSS, w, x0, st, 4242, ol.RNG_PHILOX, ol.SUM_WARP, chain_offset=9)
    r = _engine_run(engine, ol.HMC, ol.TGT_DIAG_GAUSS, w, x0, st, engine.api.ARITH_FAST, rng_mode=engine.api.RNG_PHILOX, seed=4242,
                    chain_offset=9)
    assert np.abs(r["draws"] - op).max() <= TOL
    assert np.array_equal(r["n_accept"], oap)
    assert np.all(r["draws"] >= lo) and np.all(r["draws"] <= hi)


def test_bounded_hmc_lower_only_and_l10(engine, oracle):
    """n_leap_steps = 10 is the compile-time-specialised count of the unbounded kernel; bounded runs must not take it."""
    rng = np.random.default_rng(5)
    C, d = 8, 128
    x0 = rng.uniform(0.1, 1.5, size=(C, d))
    st = ol.Settings(n_burnin=2, n_keep=20, n_leap_steps=10, step_size=0.05, lower_bounds=np.zeros(d))
    od, oa, _ = _oracle_chains(oracle, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, 55, ol.RNG_PHILOX, ol.SUM_WARP)
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = _engine_run(engine, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, arith, rng_mode=engine.api.RNG_PHILOX, seed=55)
        assert np.abs(r["draws"] - od).max() <= TOL
        assert np.array_equal(r["n_accept"], oa)
        assert r["draws"].min() > 0.0


@pytest.mark.parametrize("d", [5, 128, 300])
def test_bounded_mala_many_chains_vs_oracle(engine, oracle, d):
    rng = np.random.default_rng(200 + d)
    C = 12
    lo, hi = _mixed_bounds(d, rng)
    x0 = _start(C, d, lo, hi, rng)
    w = np.exp(rng.uniform(-0.5, 0.5, d))
    st = ol.Settings(n_burnin=3, n_keep=40, step_size=0.6 / d ** 0.25, lower_bounds=lo, upper_bounds=hi)
    od, oa, _ = _oracle_chains(oracle, ol.MALA, ol.TGT_DIAG_GAUSS, w, x0, st, 888, ol.RNG_MT, ol.SUM_WARP, mala_exact=0)
    assert 0 < oa.sum() < C * 40
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = _engine_run(engine, ol.MALA, ol.TGT_DIAG_GAUSS, w, x0, st, arith, rng_mode=engine.api.RNG_MT19937_TAPE, seed=888)
        assert np.abs(r["draws"] - od).max() <= TOL, np.abs(r["draws"] - od).max()
        assert np.array_equal(r["n_accept"], oa)
    # the literal reference (two dmvnorm() calls per draw) on the first chains
    for c in range(3):
        e = oracle.run_chain(ol.MALA, ol.TGT_DIAG_GAUSS, w, x0[c], st, seed=888 + c, sum_mode=ol.SUM_SEQ, mala_exact=1)
        assert np.abs(r["draws"][c] - e["draws"]).max() <= TOL


def test_bounded_nuts_vs_oracle(engine, oracle):
    rng = np.random.default_rng(300)
    C, d = 6, 20
    lo, hi = _mixed_bounds(d, rng)
    x0 = _start(C, d, lo, hi, rng)
    w = np.exp(rng.uniform(-0.5, 0.5, d))
    st = ol.Settings(n_burnin=0, n_keep=25, n_adapt_draws=0, step_size=0.04, lower_bounds=lo, upper_bounds=hi)
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r, od, _ = _run_pair(engine, oracle, ol.TGT_DIAG_GAUSS, "diag_gauss", w, x0, st, 999, arith)
        assert np.all(r["draws"] >= lo) and np.all(r["draws"] <= hi)
    st = ol.Settings(n_burnin=25, n_keep=25, n_adapt_draws=25, lower_bounds=lo, upper_bounds=hi)
    _run_pair(engine, oracle, ol.TGT_DIAG_GAUSS, "diag_gauss", w, x0, st, 1999, engine.api.ARITH_STRICT, tol=2e-5)


def test_bounded_rmhmc_many_chains_vs_oracle(engine, oracle):
    xs = 2 + 2 * np.sin(np.arange(100.0))
    nm = [100.0, float(xs.mean()), float(((xs - xs.mean()) ** 2).sum())]
    rng = np.random.default_rng(400)
    C = 40
    x0 = np.stack([rng.uniform(1.5, 3.0, C), rng.uniform(1.5, 3.0, C)], axis=1)
    st = ol.Settings(n_burnin=5, n_keep=40, n_leap_steps=2, step_size=0.1, lower_bounds=[-np.inf, 0.0], upper_bounds=[10.0, np.inf])
    od, oa, _ = _oracle_chains(oracle, ol.RMHMC, ol.TGT_NORMAL_MODEL, nm, x0, st, 31, ol.RNG_MT, ol.SUM_SEQ)
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = _engine_run(engine, ol.RMHMC, ol.TGT_NORMAL_MODEL, nm, x0, st, arith, rng_mode=engine.api.RNG_MT19937_TAPE, seed=31)
        assert np.abs(r["draws"] - od).max() <= 1e-9, np.abs(r["draws"] - od).max()
        assert np.array_equal(r["n_accept"], oa)
        assert r["draws"][:, :, 1].min() > 0.0


def test_bounded_hmc_at_scale_stays_inside_and_matches_sampled_chains(engine, oracle):
    """Size-independent properties on 2048 chains: every stored draw is strictly inside the box, the result does not
    depend on how the chains are sharded over calls, and sampled chains equal the oracle.  (No moment check against the
    truncated normal: with the reference's Q9 force the bounded sampler accepts rarely unless eps is tiny — the oracle
    ensemble shows the same — so parity with the reference, not the target's moments, is the meaningful bar.)"""
    C, d = 2048, 16
    rng = np.random.default_rng(8)
    x0 = rng.uniform(0.2, 1.5, size=(C, d))
    st = ol.Settings(n_burnin=20, n_keep=30, n_leap_steps=8, step_size=0.02, lower_bounds=np.zeros(d), upper_bounds=np.full(d, 3.0))
    r = _engine_run(engine, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, engine.api.ARITH_FAST, rng_mode=engine.api.RNG_PHILOX, seed=7)
    assert r["draws"].min() > 0.0 and r["draws"].max() < 3.0
    assert r["n_accept"].sum() > 0
    half = _engine_run(engine, ol.HMC, ol.TGT_ISO_GAUSS, None, x0[1024:], st, engine.api.ARITH_FAST, rng_mode=engine.api.RNG_PHILOX, seed=7,
                       chain_offset=1024)
    assert np.array_equal(half["draws"], r["draws"][1024:])
    for c in (0, 1, 777, 2047):
        o = oracle.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, x0[c], st, seed=7, rng_mode=ol.RNG_PHILOX, chain_id=c, sum_mode=ol.SUM_WARP)
        assert np.abs(r["draws"][c] - o["draws"]).max() <= TOL
        assert r["n_accept"][c] == o["n_accept"]


def test_bounds_error_paths(engine):
    api = engine.api
    x0 = np.full((2, 4), 0.5)
    with pytest.raises(api.McmcB200Error) as e:   # lower >= upper
        engine.hmc(x0, "iso_gauss", n_keep=2, n_burnin=0, lower_bounds=np.ones(4), upper_bounds=np.ones(4))
    assert e.value.code == api.ERR_INVALID_ARG
    with pytest.raises(api.McmcB200Error) as e:   # wide kernels carry no bounds
        engine.hmc(np.full((2, 1024), 0.5), "iso_gauss", n_keep=2, n_burnin=0, lower_bounds=np.zeros(1024))
    assert e.value.code == api.ERR_UNSUPPORTED


@pytest.mark.parametrize("d", [3, 6, 40, 130])
def test_bounds_together_with_a_dense_precond_mat(engine, oracle, reference, d):
    """vals_bound AND precond_mat (reachable in the reference: src/hmc.cpp:107-122, src/nuts.cpp:111-154): the kick uses J o grad
    and the drift (eps M^-1) p.  STRICT on the reference's stream and FAST on Philox against the oracle (warp order); the
    oracle itself is bit-identical to the unmodified reference on these cases.  MALA with bounds and a precond_mat is refused
    on the device (engine.cu explains why); its restatement in the oracle is still pinned to the reference here."""
    rng = np.random.default_rng(100 + d)
    C = 5
    lo, hi = _mixed_bounds(d, rng)
    x0 = _start(C, d, lo, hi, rng)
    a = rng.normal(size=(d, d))
    M = a @ a.T / d + 0.6 * np.eye(d)
    w = np.exp(rng.uniform(-0.7, 0.7, size=d))
    nk = 25 if d <= 40 else 10
    st = ol.Settings(n_burnin=3, n_keep=nk, n_leap_steps=4, step_size=0.15, precond=M, lower_bounds=lo, upper_bounds=hi)
    stm = ol.Settings(n_burnin=3, n_keep=nk, step_size=0.2, precond=M, lower_bounds=lo, upper_bounds=hi)
    if d <= 6:   # the oracle's restatement == the unmodified reference, bit for bit (bounded + dense M), HMC and MALA
        for sampler, s_ in ((ol.HMC, st), (ol.MALA, stm)):
            ref, acc, _ = reference.run_chains(sampler, ol.TGT_DIAG_GAUSS, w, x0, s_, 50)
            for c in range(C):
                o = oracle.run_chain(sampler, ol.TGT_DIAG_GAUSS, w, x0[c], s_, seed=50 + c, rng_mode=ol.RNG_MT)
                assert np.array_equal(o["draws"], ref[c]) and o["n_accept"] == acc[c], (sampler, c)
    r = _engine_run(engine, ol.HMC, ol.TGT_DIAG_GAUSS, w, x0, st, engine.api.ARITH_STRICT, rng_mode=engine.api.RNG_MT19937_TAPE, seed=50, precond_mat=M)
    for c in range(C):
        o = oracle.run_chain(ol.HMC, ol.TGT_DIAG_GAUSS, w, x0[c], st, seed=50 + c, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_WARP)
        assert np.abs(r["draws"][c] - o["draws"]).max() <= TOL, (d, c, np.abs(r["draws"][c] - o["draws"]).max())
        assert r["n_accept"][c] == o["n_accept"], (d, c)
    rf = _engine_run(engine, ol.HMC, ol.TGT_DIAG_GAUSS, w, x0, st, engine.api.ARITH_FAST, rng_mode=engine.api.RNG_PHILOX, seed=51, chain_offset=3, precond_mat=M)
    for c in range(C):
        o = oracle.run_chain(ol.HMC, ol.TGT_DIAG_GAUSS, w, x0[c], st, seed=51, rng_mode=ol.RNG_PHILOX, chain_id=3 + c, sum_mode=ol.SUM_WARP)
        assert np.abs(rf["draws"][c] - o["draws"]).max() <= TOL and rf["n_accept"][c] == o["n_accept"], (d, c)
    if d <= 40:   # NUTS, no adaptation (contract tolerance), oracle tape protocol
        stn = ol.Settings(n_burnin=2, n_keep=12, step_size=0.12, n_adapt_draws=0, max_tree_depth=6, precond=M, lower_bounds=lo, upper_bounds=hi)
        _run_pair(engine, oracle, ol.TGT_DIAG_GAUSS, "diag_gauss", w, x0, stn, 60, engine.api.ARITH_STRICT, precond=M)
    with pytest.raises(engine.McmcB200Error) as e:
        _engine_run(engine, ol.MALA, ol.TGT_DIAG_GAUSS, w, x0, stm, engine.api.ARITH_FAST, rng_mode=engine.api.RNG_PHILOX, seed=1, precond_mat=M)
    assert e.value.code == engine.api.ERR_UNSUPPORTED
