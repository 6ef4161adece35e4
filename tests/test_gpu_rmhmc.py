"""GPU parity tests for the RM-HMC path (src/rmhmc.cpp) through the C ABI: one thread per chain, metric functor of
the examples' Normal(mu, sigma) model (examples/eigen/rmhmc_normal.cpp)."""
import numpy as np
import pytest

import oracle_lib as ol
from test_gpu_hmc import TOL
from parity_util import assert_tracks_or_flips_at_threshold

pytestmark = pytest.mark.gpu


def _data():
    xs = 2 + 2 * np.sin(np.arange(100.0))
    return np.array([100.0, xs.mean(), ((xs - xs.mean()) ** 2).sum()])


def test_g4_vs_reference(engine, reference):
    """x0=(3,3), seed 1, eps=0.2, L=1, n_fp=5, keep 5 (2 accepted)."""
    td = _data()
    st = ol.Settings(n_burnin=0, n_keep=5, n_leap_steps=1, step_size=0.2)
    ref, acc = reference.run_chain(ol.RMHMC, ol.TGT_NORMAL_MODEL, td, [3, 3], st, 1)
    assert acc == 2
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = engine.rmhmc(np.array([[3.0, 3.0]]), "normal_model", target_data=td, n_leap_steps=1, step_size=0.2, n_burnin=0,
                         n_keep=5, rng_mode=engine.api.RNG_MT19937_TAPE, seed=1, arith=arith)
        assert np.abs(r["draws"][0] - ref).max() <= TOL
        assert r["n_accept"][0] == acc


@pytest.mark.parametrize("L,eps", [(2, 0.15), (3, 0.1)])
def test_many_seeds_vs_reference(engine, reference, oracle, L, eps):
    td = _data()
    C = 40
    rng = np.random.default_rng(L)
    x0 = np.array([3.0, 3.0]) + 0.2 * rng.normal(size=(C, 2))
    st = ol.Settings(n_burnin=10, n_keep=150, n_leap_steps=L, step_size=eps)
    ref, acc, _ = reference.run_chains(ol.RMHMC, ol.TGT_NORMAL_MODEL, td, x0, st, 500)
    r = engine.rmhmc(x0, "normal_model", target_data=td, n_leap_steps=L, step_size=eps, n_burnin=10, n_keep=150,
                     rng_mode=engine.api.RNG_MT19937_TAPE, seed=500, arith=engine.api.ARITH_STRICT, want_logp=True)
    # the Normal model is nonlinear: a rounding-level difference (device log vs glibc log) can in principle flip an
    # accept decision.  Every chain must track the oracle (== the reference, bit for bit) to the contract tolerance, or leave
    # its path at a draw whose accept margin |u - exp(comp)| is at rounding level (parity_util) — nothing else passes.
    tracked = 0
    for c in range(C):
        o = oracle.run_chain(ol.RMHMC, ol.TGT_NORMAL_MODEL, td, x0[c], st, seed=500 + c, rng_mode=ol.RNG_MT, want_margins=True)
        assert np.array_equal(o["draws"], ref[c]) and o["n_accept"] == acc[c]   # oracle == unmodified reference
        if assert_tracks_or_flips_at_threshold(r["draws"][c], o, 10, TOL, "chain %d" % c):
            assert r["n_accept"][c] == acc[c]
            tracked += 1
    assert tracked >= C - 1, tracked
    assert 0 < acc.min() and acc.max() < 150


def test_philox_mode_vs_oracle(engine, oracle):
    td = _data()
    C = 16
    x0 = np.tile([3.0, 3.0], (C, 1))
    st = ol.Settings(n_burnin=5, n_keep=80, n_leap_steps=2, step_size=0.15)
    r = engine.rmhmc(x0, "normal_model", target_data=td, n_leap_steps=2, step_size=0.15, n_burnin=5, n_keep=80,
                     rng_mode=engine.api.RNG_PHILOX, seed=31337, chain_offset=7)
    for c in range(C):
        o = oracle.run_chain(ol.RMHMC, ol.TGT_NORMAL_MODEL, td, x0[c], st, seed=31337, rng_mode=ol.RNG_PHILOX, chain_id=7 + c, want_margins=True)
        if assert_tracks_or_flips_at_threshold(r["draws"][c], o, 5, TOL, "chain %d" % c):
            assert r["n_accept"][c] == o["n_accept"]
    assert np.abs(r["draws"][0] - r["draws"][1]).max() > 1e-6  # distinct substreams


def test_moments_track_reference_and_unsupported_target(engine, reference):
    """4096 Philox chains vs 400 reference chains (own mt19937 streams), same settings: the ensemble moments after the
    same number of draws must agree statistically (the reference's sampler is bug-compatible, not exact: Q16/Q17)."""
    td = _data()
    C = 4096
    x0 = np.tile([3.0, 3.0], (C, 1))
    r = engine.rmhmc(x0, "normal_model", target_data=td, n_leap_steps=2, step_size=0.2, n_burnin=200, n_keep=100,
                     rng_mode=engine.api.RNG_PHILOX, seed=1)
    st = ol.Settings(n_burnin=200, n_keep=100, n_leap_steps=2, step_size=0.2)
    ref, racc, _ = reference.run_chains(ol.RMHMC, ol.TGT_NORMAL_MODEL, td, x0[:400], st, 1)
    m, mr = r["draws"].mean(axis=(0, 1)), ref.mean(axis=(0, 1))
    assert np.abs(m - mr).max() < 0.06, (m, mr)
    acc, acc_r = r["n_accept"].mean() / 100, racc.mean() / 100
    assert abs(acc - acc_r) < 0.03, (acc, acc_r)
    with pytest.raises(engine.McmcB200Error) as ei:
        engine.rmhmc(np.zeros((2, 4)), "iso_gauss", n_burnin=1, n_keep=1)
    assert ei.value.code == engine.api.ERR_UNSUPPORTED


# ---- the warp-per-chain kernel for general n_dim <= 64 (rmhmc_general.cu) and Neal's funnel (BASELINE config 5) -------

def test_general_kernel_reproduces_thread_per_chain_kernel_and_goldens(engine, reference, monkeypatch):
    """MCMCB200_RMHMC_GENERAL=1 routes the 2-parameter Normal model through the general kernel (LU inverse, Cholesky and
    log-det on matrices in global scratch): it must give what the register kernel gives, and the reference's G4 golden."""
    td = _data()
    rng = np.random.default_rng(4)
    x0 = np.array([3.0, 3.0]) + 0.2 * rng.normal(size=(24, 2))
    kw = dict(target_data=td, n_leap_steps=2, step_size=0.15, n_burnin=5, n_keep=60, want_logp=True)
    for mode, seed in ((engine.api.RNG_MT19937_TAPE, 77), (engine.api.RNG_PHILOX, 78)):
        for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
            a = engine.rmhmc(x0, "normal_model", rng_mode=mode, seed=seed, arith=arith, **kw)
            monkeypatch.setenv("MCMCB200_RMHMC_GENERAL", "1")
            b = engine.rmhmc(x0, "normal_model", rng_mode=mode, seed=seed, arith=arith, **kw)
            monkeypatch.delenv("MCMCB200_RMHMC_GENERAL")
            if arith == engine.api.ARITH_STRICT:
                assert np.array_equal(a["draws"], b["draws"]) and np.array_equal(a["logp"], b["logp"])
            else:
                assert np.abs(a["draws"] - b["draws"]).max() <= TOL
            assert np.array_equal(a["n_accept"], b["n_accept"])
    st = ol.Settings(n_burnin=0, n_keep=5, n_leap_steps=1, step_size=0.2)
    ref, acc = reference.run_chain(ol.RMHMC, ol.TGT_NORMAL_MODEL, td, [3, 3], st, 1)
    monkeypatch.setenv("MCMCB200_RMHMC_GENERAL", "1")
    r = engine.rmhmc(np.array([[3.0, 3.0]]), "normal_model", target_data=td, n_leap_steps=1, step_size=0.2, n_burnin=0, n_keep=5,
                     rng_mode=engine.api.RNG_MT19937_TAPE, seed=1, arith=engine.api.ARITH_STRICT)
    assert np.abs(r["draws"][0] - ref).max() <= TOL and r["n_accept"][0] == acc


def _close(a, b, tol=TOL):
    """L-inf agreement where both are finite, identical NaN pattern elsewhere: the reference ACCEPTS a proposal whose energy
    is NaN (std::min(0.01, NaN) = 0.01, src/rmhmc.cpp:250), so an unstable chain turns NaN — in the reference, the oracle and
    the kernel at the same draw."""
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a, b, rtol=0, atol=tol, equal_nan=True)


def _funnel_start(C, d, rng):
    x0 = rng.normal(size=(C, d)) * 0.6
    x0[:, 0] = rng.uniform(-0.5, 0.8, size=C)
    return x0


@pytest.mark.parametrize("metric_id", [1, 2])
@pytest.mark.parametrize("d,L,eps", [(2, 2, 0.1), (3, 3, 0.1), (5, 2, 0.15), (8, 3, 0.08), (17, 2, 0.1), (33, 2, 0.08), (64, 2, 0.06)])
def test_funnel_rmhmc_vs_oracle_and_reference(engine, oracle, reference, d, L, eps, metric_id):
    """C5-shaped runs: Neal's funnel with its registered position-dependent metric, general n_dim up to 64.  STRICT on the
    reference's stream and FAST on Philox against the oracle (SUM_WARP), and against the unmodified reference where its
    O(d^4) momentum updates are affordable.  exp/log/sqrt are the CUDA library's, so the tolerance is 1e-10, not bits; as
    for the Normal model a rounding-level difference may flip one accept decision and decorrelate one chain.
    metric_id 1 = the Fisher-type diagonal metric, 2 = the closed-form SoftAbs metric (alpha = 1e6) of BASELINE config 5."""
    rng = np.random.default_rng(d)
    C = 6
    if metric_id == 2 and d >= 17:
        eps = 0.02   # the reference's RM-HMC (Q16/Q17) with SoftAbs needs small steps at this size to accept anything
    x0 = _funnel_start(C, d, rng)
    nk = 30 if d <= 17 else 12
    st = ol.Settings(n_burnin=2, n_keep=nk, n_leap_steps=L, step_size=eps, n_fp_steps=4, metric_id=metric_id)
    kw = dict(n_leap_steps=L, step_size=eps, n_fp_steps=4, n_burnin=2, n_keep=nk, want_logp=True, metric_id=metric_id)
    r = engine.rmhmc(x0, "funnel", rng_mode=engine.api.RNG_MT19937_TAPE, seed=900, arith=engine.api.ARITH_STRICT, **kw)
    tracked = []
    for c in range(C):
        o = oracle.run_chain(ol.RMHMC, ol.TGT_FUNNEL, None, x0[c], st, seed=900 + c, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_WARP, want_logp=True,
                             want_margins=True)
        pert = lambda c=c: oracle.run_chain(ol.RMHMC, ol.TGT_FUNNEL, None, np.nextafter(x0[c], np.inf), st, seed=900 + c, rng_mode=ol.RNG_MT,
                                            sum_mode=ol.SUM_WARP)["draws"]
        ok = assert_tracks_or_flips_at_threshold(r["draws"][c], o, 2, TOL, "strict d=%d chain %d" % (d, c), rerun_perturbed=pert)
        if ok:
            assert r["n_accept"][c] == o["n_accept"] and _close(r["logp"][c], o["logp"], 1e-9)
        tracked.append(ok)
    assert sum(tracked) >= C - 1
    assert 0 < r["n_accept"].sum()
    if d <= 8:
        ref, acc, _ = reference.run_chains(ol.RMHMC, ol.TGT_FUNNEL, None, x0, st, 900)
        for c in range(C):
            if tracked[c]:
                assert _close(r["draws"][c], ref[c]) and r["n_accept"][c] == acc[c]
    rf = engine.rmhmc(x0, "funnel", rng_mode=engine.api.RNG_PHILOX, seed=901, chain_offset=11, arith=engine.api.ARITH_FAST, **kw)
    n_fast_tracked = 0
    for c in range(C):
        o = oracle.run_chain(ol.RMHMC, ol.TGT_FUNNEL, None, x0[c], st, seed=901, rng_mode=ol.RNG_PHILOX, chain_id=11 + c, sum_mode=ol.SUM_WARP,
                             want_margins=True)
        pert = lambda c=c: oracle.run_chain(ol.RMHMC, ol.TGT_FUNNEL, None, np.nextafter(x0[c], np.inf), st, seed=901, rng_mode=ol.RNG_PHILOX,
                                            chain_id=11 + c, sum_mode=ol.SUM_WARP)["draws"]
        if assert_tracks_or_flips_at_threshold(rf["draws"][c], o, 2, TOL, "fast d=%d chain %d" % (d, c), rerun_perturbed=pert):
            assert rf["n_accept"][c] == o["n_accept"]
            n_fast_tracked += 1
    assert n_fast_tracked >= C - 2, n_fast_tracked


def test_funnel_other_samplers_and_unsupported_combinations(engine, oracle):
    """The funnel functor under HMC / MALA / RWMH / NUTS (Philox, FAST) against the oracle; RM-HMC refuses n_dim > 64."""
    rng = np.random.default_rng(12)
    for d in (2, 9, 64, 100):
        C = 5
        x0 = _funnel_start(C, d, rng)
        for smp, call, st in (
                (ol.HMC, lambda: engine.hmc(x0, "funnel", n_leap_steps=4, step_size=0.05, n_burnin=2, n_keep=20, rng_mode=engine.api.RNG_PHILOX, seed=5),
                 ol.Settings(n_burnin=2, n_keep=20, n_leap_steps=4, step_size=0.05)),
                (ol.MALA, lambda: engine.mala(x0, "funnel", step_size=0.05, n_burnin=2, n_keep=20, rng_mode=engine.api.RNG_PHILOX, seed=5),
                 ol.Settings(n_burnin=2, n_keep=20, step_size=0.05)),
                (ol.RWMH, lambda: engine.rwmh(x0, "funnel", par_scale=0.05, n_burnin=2, n_keep=20, rng_mode=engine.api.RNG_PHILOX, seed=5),
                 ol.Settings(n_burnin=2, n_keep=20, step_size=0.05)),
                (ol.NUTS, lambda: engine.nuts(x0, "funnel", step_size=0.05, n_adapt_draws=0, max_tree_depth=6, n_burnin=2, n_keep=8,
                                              rng_mode=engine.api.RNG_PHILOX, seed=5),
                 ol.Settings(n_burnin=2, n_keep=8, step_size=0.05, n_adapt_draws=0, max_tree_depth=6))):
            r = call()
            for c in range(C):
                o = oracle.run_chain(smp, ol.TGT_FUNNEL, None, x0[c], st, seed=5, rng_mode=ol.RNG_PHILOX, chain_id=c, sum_mode=ol.SUM_WARP)
                assert np.abs(r["draws"][c] - o["draws"]).max() <= TOL and r["n_accept"][c] == o["n_accept"], (smp, d, c)
    with pytest.raises(engine.McmcB200Error) as ei:   # beyond the CTA kernel's generic mappings (n_dim <= 128)
        engine.rmhmc(np.zeros((2, 129)), "funnel", n_burnin=1, n_keep=1)
    assert ei.value.code == engine.api.ERR_UNSUPPORTED
    with pytest.raises(engine.McmcB200Error) as ei:   # STRICT arithmetic runs on the cube kernel: n_dim <= 64
        engine.rmhmc(np.zeros((2, 65)), "funnel", n_burnin=1, n_keep=1, arith=engine.api.ARITH_STRICT)
    assert ei.value.code == engine.api.ERR_UNSUPPORTED
    with pytest.raises(engine.McmcB200Error) as ei:
        engine.rmhmc(np.zeros((2, 8)), "funnel", n_burnin=1, n_keep=1, metric_id=7)
    assert ei.value.code == engine.api.ERR_UNSUPPORTED


def test_funnel_goldens(engine):
    """Committed fixtures from the unmodified reference: rmhmc_funnel_{fisher,softabs}_d6 and hmc_funnel_d6."""
    import golden_util

    seen = 0
    for c in golden_util.load()["cases"]:
        if c["target"] != ol.TGT_FUNNEL:
            continue
        st = c["settings"]
        x0 = np.array([c["x0"]], dtype=np.float64)
        common = dict(n_burnin=st["n_burnin"], n_keep=st["n_keep"], rng_mode=engine.api.RNG_MT19937_TAPE, seed=c["seed"], arith=engine.api.ARITH_STRICT)
        if c["sampler"] == ol.RMHMC:
            r = engine.rmhmc(x0, "funnel", n_leap_steps=st["n_leap_steps"], step_size=st["step_size"], n_fp_steps=st["n_fp_steps"],
                             metric_id=st["metric_id"], **common)
        else:
            r = engine.hmc(x0, "funnel", n_leap_steps=st["n_leap_steps"], step_size=st["step_size"], **common)
        assert np.abs(r["draws"][0] - c["draws"]).max() <= TOL, c["name"]
        assert r["n_accept"][0] == c["n_accept"], c["name"]
        seen += 1
    assert seen == 3


@pytest.mark.parametrize("metric_id", [1, 2])
def test_cta_kernel_matches_the_cube_kernel(engine, monkeypatch, metric_id):
    """FAST arithmetic runs one CTA per chain with the metric algebra in shared memory and the derivative cube replaced by the
    metric's closed-form contractions (rmhmc_cta.cu); MCMCB200_RMHMC_CTA=0 forces the warp-per-chain cube kernel
    (rmhmc_general.cu).  Same algorithm, different operation order: <= 1e-10 and identical accept counts, Philox and tape,
    ragged and full n_dim, both chol modes."""
    rng = np.random.default_rng(50 + metric_id)
    for d, L, eps in ((2, 2, 0.1), (7, 3, 0.08), (33, 2, 0.05), (64, 3, 0.02)):
        x0 = _funnel_start(12, d, rng)
        for mode, chol in ((engine.api.RNG_PHILOX, 1), (engine.api.RNG_MT19937_TAPE, 0)):
            kw = dict(n_leap_steps=L, step_size=eps, n_fp_steps=4, n_burnin=2, n_keep=10, want_logp=True, metric_id=metric_id, rng_mode=mode, seed=77,
                      chol_mode=chol)
            a = engine.rmhmc(x0, "funnel", **kw)
            monkeypatch.setenv("MCMCB200_RMHMC_CTA", "0")
            b = engine.rmhmc(x0, "funnel", **kw)
            b2 = engine.rmhmc(np.nextafter(x0, np.inf), "funnel", **kw)   # the cube kernel again, start moved by one ulp
            monkeypatch.delenv("MCMCB200_RMHMC_CTA")
            # chains on which the reference algorithm is numerically stable (see parity_util): the cube kernel reproduces itself
            # from a one-ulp-perturbed start.  The others (the reference accepts NaN energies; large steps diverge) carry no
            # information about operation order.
            blown = lambda r_, c: (not np.isfinite(r_["draws"][c]).all()) or np.abs(r_["draws"][c]).max() > 1e30   # overflowed trajectories:
            stable = np.array([_close(b["draws"][c], b2["draws"][c], 0.25 * TOL) and not blown(a, c) and not blown(b, c)   # inf / NaN arithmetic in
                               for c in range(x0.shape[0])])                                                              # the accept test is implementation-defined
            assert stable.sum() >= 7, (d, mode, stable)
            assert _close(a["draws"][stable], b["draws"][stable]), (d, mode, np.nanmax(np.abs(a["draws"][stable] - b["draws"][stable])))
            assert np.array_equal(a["n_accept"][stable], b["n_accept"][stable]) and _close(a["logp"][stable], b["logp"][stable], 1e-8)
            assert a["n_accept"].sum() > 0


def test_c5_full_size(engine, monkeypatch):
    """BASELINE config 5 at full size: RM-HMC, Neal's funnel d = 64, SoftAbs metric, 2048 chains, L = 5, n_fp = 5 (a few draws).
    Every chain stays finite, the acceptance rate is in the range the reference's sampler shows at this step size, and a
    subset of the chains agrees with the cube kernel draw for draw."""
    rng = np.random.default_rng(5)
    d, C = 64, 2048
    x0 = _funnel_start(C, d, rng)
    kw = dict(n_leap_steps=5, step_size=0.01, n_fp_steps=5, n_burnin=1, n_keep=5, rng_mode=engine.api.RNG_PHILOX, seed=5, metric_id=2)
    r = engine.rmhmc(x0, "funnel", **kw)
    assert np.isfinite(r["draws"]).all()
    acc = r["n_accept"].mean() / 5
    assert 0.2 < acc < 0.95, acc
    monkeypatch.setenv("MCMCB200_RMHMC_CTA", "0")
    sub = engine.rmhmc(x0[1000:1024], "funnel", chain_offset=1000, **kw)
    monkeypatch.delenv("MCMCB200_RMHMC_CTA")
    assert np.abs(sub["draws"] - r["draws"][1000:1024]).max() <= TOL and np.array_equal(sub["n_accept"], r["n_accept"][1000:1024])
    print("C5 full size: kernel %.1f ms for 6 draws (%.2f ms/draw), acceptance %.2f" % (r["kernel_ms"], r["kernel_ms"] / 6, acc))


def test_register_tile_elimination_reproduces_the_shared_memory_one(engine, monkeypatch):
    """n_dim > 32: the CTA kernel inverts its metrics with the matrix held in registers (8 x 4 tile per thread, implicit
    partial pivoting, rc_inverse_regtile); MCMCB200_RMHMC_REGTILE=0 keeps the shared-memory Gauss-Jordan with physical row
    exchanges.  Same pivots, same multipliers, same update formulas: identical bits, full and ragged n_dim, both metrics."""
    rng = np.random.default_rng(77)
    for d, metric_id in ((64, 2), (40, 2), (33, 1), (63, 2)):
        x0 = _funnel_start(24, d, rng)
        kw = dict(n_leap_steps=3, step_size=0.02, n_fp_steps=4, n_burnin=1, n_keep=6, want_logp=True, metric_id=metric_id,
                  rng_mode=engine.api.RNG_PHILOX, seed=9)
        a = engine.rmhmc(x0, "funnel", **kw)
        monkeypatch.setenv("MCMCB200_RMHMC_REGTILE", "0")
        b = engine.rmhmc(x0, "funnel", **kw)
        monkeypatch.delenv("MCMCB200_RMHMC_REGTILE")
        assert np.array_equal(a["draws"], b["draws"], equal_nan=True), (d, metric_id, np.nanmax(np.abs(a["draws"] - b["draws"])))
        assert np.array_equal(a["n_accept"], b["n_accept"]) and np.array_equal(a["logp"], b["logp"], equal_nan=True)


def test_generic_thread_mappings_and_n_dim_up_to_128(engine, oracle, monkeypatch):
    """64 < n_dim <= 128: the CTA kernel switches from its tuned mappings (two threads per row, register-tile elimination) to
    generic ones (one thread per row, 64 row pairs x 2 column groups in the elimination; the 133 KB matrix leaves one CTA per
    SM).  (i) MCMCB200_RMHMC_WIDE=1 forces the generic mappings at n_dim <= 64, where they must reproduce the tuned ones bit for
    bit — every element sees the same operations in the same order; (ii) n_dim = 96 against the oracle (Philox, FAST), both
    funnel metrics; (iii) n_dim = 128 stays finite and accepts."""
    rng = np.random.default_rng(123)
    for d, metric_id in ((64, 2), (40, 1), (7, 2), (33, 2)):
        x0 = _funnel_start(16, d, rng)
        kw = dict(n_leap_steps=3, step_size=0.02, n_fp_steps=4, n_burnin=1, n_keep=5, want_logp=True, metric_id=metric_id,
                  rng_mode=engine.api.RNG_PHILOX, seed=19)
        a = engine.rmhmc(x0, "funnel", **kw)
        monkeypatch.setenv("MCMCB200_RMHMC_WIDE", "1")
        b = engine.rmhmc(x0, "funnel", **kw)
        monkeypatch.delenv("MCMCB200_RMHMC_WIDE")
        assert np.array_equal(a["draws"], b["draws"], equal_nan=True), (d, metric_id, np.nanmax(np.abs(a["draws"] - b["draws"])))
        assert np.array_equal(a["n_accept"], b["n_accept"]) and np.array_equal(a["logp"], b["logp"], equal_nan=True)
    d, C = 96, 3
    for metric_id in (2, 1):
        x0 = _funnel_start(C, d, rng)
        st = ol.Settings(n_burnin=1, n_keep=4, n_leap_steps=2, step_size=0.02, n_fp_steps=3, metric_id=metric_id)
        r = engine.rmhmc(x0, "funnel", n_leap_steps=2, step_size=0.02, n_fp_steps=3, n_burnin=1, n_keep=4, metric_id=metric_id,
                         rng_mode=engine.api.RNG_PHILOX, seed=77, chain_offset=5)
        tracked = 0
        for c in range(C):
            o = oracle.run_chain(ol.RMHMC, ol.TGT_FUNNEL, None, x0[c], st, seed=77, rng_mode=ol.RNG_PHILOX, chain_id=5 + c, sum_mode=ol.SUM_WARP,
                                 want_margins=True)
            pert = lambda c=c: oracle.run_chain(ol.RMHMC, ol.TGT_FUNNEL, None, np.nextafter(x0[c], np.inf), st, seed=77, rng_mode=ol.RNG_PHILOX,
                                                chain_id=5 + c, sum_mode=ol.SUM_WARP)["draws"]
            if assert_tracks_or_flips_at_threshold(r["draws"][c], o, 1, TOL, "n_dim=96 metric %d chain %d" % (metric_id, c), rerun_perturbed=pert):
                assert r["n_accept"][c] == o["n_accept"]
                tracked += 1
        assert tracked >= C - 1, (metric_id, tracked)
    x0 = _funnel_start(8, 128, rng)
    r = engine.rmhmc(x0, "funnel", n_leap_steps=2, step_size=0.01, n_fp_steps=3, n_burnin=1, n_keep=4, metric_id=2, rng_mode=engine.api.RNG_PHILOX, seed=3)
    assert np.isfinite(r["draws"]).all() and r["n_accept"].sum() > 0
