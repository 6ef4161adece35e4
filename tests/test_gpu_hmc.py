"""GPU parity tests for the HMC path: CUDA kernels (through the C ABI) vs the CPU oracle and the
unmodified reference.  Tolerance of the contract (BASELINE.json north_star): per-draw L-inf <= 1e-10 and
|delta log pi| <= 1e-10 in fp64.  STRICT arithmetic is additionally required to be bit-exact against the
oracle run with the kernels' reduction order."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _oracle_chains(oracle, sampler, tid, tdata, x0s, st, seed, rng_mode, sum_mode, chain_offset=0, **kw):
    draws, acc, logp = [], [], []
    for c in range(x0s.shape[0]):
        s = seed + chain_offset + c if rng_mode == ol.RNG_MT else seed
        r = oracle.run_chain(sampler, tid, tdata, x0s[c], st, seed=s, rng_mode=rng_mode, chain_id=chain_offset + c,
                             sum_mode=sum_mode, want_logp=True, **kw)
        draws.append(r["draws"]); acc.append(r["n_accept"]); logp.append(r["logp"])
    return np.stack(draws), np.array(acc), np.stack(logp)


def test_c1_plumbing_d3(engine, reference):
    """BASELINE config 1 / SURVEY Appendix B G2: d=3 standard Gaussian, 1 chain, seed 1."""
    st = ol.Settings(n_burnin=0, n_keep=5, n_leap_steps=10, step_size=0.1)
    ref, acc = reference.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, [1, -1, 0.5], st, 1)
    for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = engine.hmc(np.array([[1, -1, 0.5]]), "iso_gauss", n_leap_steps=10, step_size=0.1, n_burnin=0, n_keep=5,
                       rng_mode=engine.api.RNG_MT19937_TAPE, seed=1, arith=arith)
        assert np.abs(r["draws"][0] - ref).max() <= TOL
        assert r["n_accept"][0] == acc
    g2 = np.array([0.21394863357965233, 0.038869637751197325, -0.40013418025888436])
    assert np.abs(r["draws"][0][0] - g2).max() <= TOL


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_c2_parity_subset_vs_reference(engine, reference, oracle, arith):
    """SURVEY §8(d) C2 parity subset: d=128 iso-Gaussian, 64 chains x 50 draws, L=10, eps=0.1, seeds 12345+c."""
    C, d = 64, 128
    x0 = ol.c2_initial(C, d)
    st = ol.Settings(n_burnin=10, n_keep=50, n_leap_steps=10, step_size=0.1)
    ref, acc, _ = reference.run_chains(ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, 12345)
    a = engine.api.ARITH_STRICT if arith == "strict" else engine.api.ARITH_FAST
    r = engine.hmc(x0, "iso_gauss", n_leap_steps=10, step_size=0.1, n_burnin=10, n_keep=50,
                   rng_mode=engine.api.RNG_MT19937_TAPE, seed=12345, arith=a, want_logp=True)
    linf = np.abs(r["draws"] - ref).max(axis=(1, 2))
    assert linf.max() <= TOL, linf.max()
    assert np.array_equal(r["n_accept"], acc)
    if arith == "strict":
        # element-wise bit-exact against the unmodified reference (only accept decisions could differ, none do)
        assert np.array_equal(r["draws"], ref)
        od, oa, olp = _oracle_chains(oracle, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, 12345, ol.RNG_MT, ol.SUM_WARP)
        assert np.array_equal(r["draws"], od)
        assert np.array_equal(r["logp"], olp)  # same reduction order -> bit-exact log pi


def test_philox_mode_vs_oracle(engine, oracle):
    C, d = 16, 128
    x0 = ol.c2_initial(C, d)
    st = ol.Settings(n_burnin=5, n_keep=40, n_leap_steps=10, step_size=0.1)
    od, oa, olp = _oracle_chains(oracle, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, 987654321012345, ol.RNG_PHILOX, ol.SUM_WARP,
                                 chain_offset=100)
    for a in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
        r = engine.hmc(x0, "iso_gauss", n_leap_steps=10, step_size=0.1, n_burnin=5, n_keep=40,
                       rng_mode=engine.api.RNG_PHILOX, seed=987654321012345, chain_offset=100, arith=a, want_logp=True)
        assert np.abs(r["draws"] - od).max() <= TOL
        assert np.abs(r["logp"] - olp).max() <= TOL
        assert np.array_equal(r["n_accept"], oa)


def test_philox_raw_stream_vs_oracle(engine, oracle):
    for d in (3, 64, 128, 200):
        for chain, draw in ((0, -1), (5, 0), (4095, 1099)):
            dev = engine.api.philox_stream(12345, chain, draw, d, 3)
            host = oracle.rng_stream(ol.RNG_PHILOX, 12345, chain, draw, d, 3)
            assert np.abs(dev - host).max() <= 1e-14
            assert np.array_equal(dev[d:], host[d:])  # uniforms are exact integer arithmetic


@pytest.mark.parametrize("d", [1, 2, 3, 5, 63, 64, 65, 70, 129, 200, 256, 257, 512])
def test_ragged_dims_strict_bitexact(engine, oracle, d):
    """Dimensions that are not multiples of the 64-element lane-pair stripe, incl. the largest supported."""
    C = 5
    rng = np.random.default_rng(d)
    x0 = rng.normal(size=(C, d))
    st = ol.Settings(n_burnin=3, n_keep=12, n_leap_steps=4, step_size=0.2)
    od, oa, olp = _oracle_chains(oracle, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, 77, ol.RNG_MT, ol.SUM_WARP)
    r = engine.hmc(x0, "iso_gauss", n_leap_steps=4, step_size=0.2, n_burnin=3, n_keep=12,
                   rng_mode=engine.api.RNG_MT19937_TAPE, seed=77, arith=engine.api.ARITH_STRICT, want_logp=True)
    assert np.array_equal(r["draws"], od)
    assert np.array_equal(r["n_accept"], oa)
    assert np.array_equal(r["logp"], olp)


def _sym_pd(rng, d, shift):
    a = rng.normal(size=(d, d))
    m = a @ a.T / d + shift * np.eye(d)
    return (m + m.T) / 2


@pytest.mark.parametrize("d", [6, 70])
@pytest.mark.parametrize("chol_mode", [0, 1])
def test_dense_mass_and_dense_target(engine, oracle, reference, d, chol_mode):
    """precond_mat path (src/hmc.cpp:57-59) incl. the Eigen matrixLLT storage quirk (Q8), dense-precision target."""
    rng = np.random.default_rng(100 + d)
    P = _sym_pd(rng, d, 1.0)
    M = _sym_pd(rng, d, 0.5)
    C = 4
    x0 = rng.normal(size=(C, d))
    st = ol.Settings(n_burnin=2, n_keep=25, n_leap_steps=5, step_size=0.15, precond=M)
    od, oa, olp = _oracle_chains(oracle, ol.HMC, ol.TGT_DENSE_GAUSS, P.ravel(), x0, st, 5, ol.RNG_MT, ol.SUM_WARP,
                                 chol_mode=chol_mode)
    r = engine.hmc(x0, "dense_gauss", target_data=P, n_leap_steps=5, step_size=0.15, precond_mat=M, n_burnin=2, n_keep=25,
                   rng_mode=engine.api.RNG_MT19937_TAPE, seed=5, arith=engine.api.ARITH_STRICT, chol_mode=chol_mode,
                   want_logp=True)
    assert np.abs(r["draws"] - od).max() <= TOL
    assert np.array_equal(r["n_accept"], oa)
    assert np.abs(r["logp"] - olp).max() <= 1e-9 * max(1.0, np.abs(olp).max())
    rf = engine.hmc(x0, "dense_gauss", target_data=P, n_leap_steps=5, step_size=0.15, precond_mat=M, n_burnin=2, n_keep=25,
                    rng_mode=engine.api.RNG_MT19937_TAPE, seed=5, arith=engine.api.ARITH_FAST, chol_mode=chol_mode)
    assert np.abs(rf["draws"] - od).max() <= TOL
    if chol_mode == 1:  # the Eigen-backend reference
        ref, acc, _ = reference.run_chains(ol.HMC, ol.TGT_DENSE_GAUSS, P.ravel(), x0, st, 5)
        assert np.abs(r["draws"] - ref).max() <= TOL
        assert np.array_equal(r["n_accept"], acc)


def test_other_targets(engine, oracle):
    rng = np.random.default_rng(3)
    # diagonal Gaussian
    d, C = 40, 6
    w = np.linspace(0.5, 3.0, d)
    x0 = rng.normal(size=(C, d))
    st = ol.Settings(n_burnin=5, n_keep=30, n_leap_steps=6, step_size=0.2)
    od, oa, _ = _oracle_chains(oracle, ol.HMC, ol.TGT_DIAG_GAUSS, w, x0, st, 9, ol.RNG_MT, ol.SUM_WARP)
    r = engine.hmc(x0, "diag_gauss", target_data=w, n_leap_steps=6, step_size=0.2, n_burnin=5, n_keep=30,
                   rng_mode=engine.api.RNG_MT19937_TAPE, seed=9, arith=engine.api.ARITH_STRICT)
    assert np.array_equal(r["draws"], od) and np.array_equal(r["n_accept"], oa)
    # linear regression posterior
    d = 24
    A = _sym_pd(rng, d, 2.0); b = rng.normal(size=d)
    x0 = rng.normal(size=(C, d))
    td = np.concatenate([A.ravel(), b])
    st = ol.Settings(n_burnin=5, n_keep=30, n_leap_steps=6, step_size=0.1)
    od, oa, _ = _oracle_chains(oracle, ol.HMC, ol.TGT_LINREG, td, x0, st, 10, ol.RNG_MT, ol.SUM_WARP)
    r = engine.hmc(x0, "linreg", target_data=td, n_leap_steps=6, step_size=0.1, n_burnin=5, n_keep=30,
                   rng_mode=engine.api.RNG_MT19937_TAPE, seed=10, arith=engine.api.ARITH_STRICT)
    assert np.abs(r["draws"] - od).max() <= TOL and np.array_equal(r["n_accept"], oa)
    # the examples' Normal(mu, sigma) model (rejections happen here)
    xs = 2 + 2 * np.sin(np.arange(100.0))
    td = np.array([100.0, xs.mean(), ((xs - xs.mean()) ** 2).sum()])
    x0 = np.tile([3.0, 3.0], (C, 1)) + 0.1 * rng.normal(size=(C, 2))
    st = ol.Settings(n_burnin=10, n_keep=100, n_leap_steps=5, step_size=0.08)
    od, oa, _ = _oracle_chains(oracle, ol.HMC, ol.TGT_NORMAL_MODEL, td, x0, st, 11, ol.RNG_MT, ol.SUM_WARP)
    r = engine.hmc(x0, "normal_model", target_data=td, n_leap_steps=5, step_size=0.08, n_burnin=10, n_keep=100,
                   rng_mode=engine.api.RNG_MT19937_TAPE, seed=11, arith=engine.api.ARITH_STRICT)
    assert np.abs(r["draws"] - od).max() <= TOL and np.array_equal(r["n_accept"], oa)
    assert 0 < oa.min() and oa.max() < 100 or oa.max() <= 100


def test_rejections_and_nonfinite(engine, oracle):
    """Large step size: many rejections; accept decisions and kept rows must track the oracle exactly."""
    C, d = 8, 32
    x0 = ol.c2_initial(C, d)
    st = ol.Settings(n_burnin=0, n_keep=200, n_leap_steps=3, step_size=1.3)
    od, oa, _ = _oracle_chains(oracle, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, 21, ol.RNG_MT, ol.SUM_WARP)
    r = engine.hmc(x0, "iso_gauss", n_leap_steps=3, step_size=1.3, n_burnin=0, n_keep=200,
                   rng_mode=engine.api.RNG_MT19937_TAPE, seed=21, arith=engine.api.ARITH_STRICT)
    assert np.array_equal(r["draws"], od) and np.array_equal(r["n_accept"], oa)
    assert oa.max() < 200  # the case really exercises rejection
    # unstable integrator (eps > 2): energy overflows to inf/nan -> reject path (src/hmc.cpp:180-182)
    st = ol.Settings(n_burnin=0, n_keep=20, n_leap_steps=400, step_size=2.5)
    od, oa, _ = _oracle_chains(oracle, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, 22, ol.RNG_MT, ol.SUM_WARP)
    r = engine.hmc(x0, "iso_gauss", n_leap_steps=400, step_size=2.5, n_burnin=0, n_keep=20,
                   rng_mode=engine.api.RNG_MT19937_TAPE, seed=22, arith=engine.api.ARITH_STRICT)
    assert np.array_equal(r["draws"], od) and np.array_equal(r["n_accept"], oa)


def test_user_tape_and_broadcast_initial(engine, oracle):
    d, C, n = 10, 3, 6
    rng = np.random.default_rng(0)
    tape = np.concatenate([rng.normal(size=(C, n, d)), rng.uniform(size=(C, n, 1))], axis=2).reshape(C, -1)
    x0 = rng.normal(size=d)
    st = ol.Settings(n_burnin=0, n_keep=n, n_leap_steps=2, step_size=0.3)
    r = engine.hmc(x0, "iso_gauss", n_leap_steps=2, step_size=0.3, n_burnin=0, n_keep=n, n_chains=C,
                   rng_mode=engine.api.RNG_USER_TAPE, tape=tape, arith=engine.api.ARITH_STRICT)
    for c in range(C):
        o = oracle.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, rng_mode=ol.RNG_TAPE, tape=tape[c], sum_mode=ol.SUM_WARP)
        assert np.array_equal(r["draws"][c], o["draws"])


def test_full_size_c2_properties(engine):
    """BASELINE config 2 at full size (4096 chains, d=128, L=10, eps=0.1, 100+1000 draws) through size-independent
    properties: sharding invariance (chain_offset), stationarity moments of N(0, I), acceptance rate, finiteness."""
    C, d = 4096, 128
    x0 = ol.c2_initial(C, d)
    kw = dict(n_leap_steps=10, step_size=0.1, n_burnin=100, n_keep=1000, rng_mode=engine.api.RNG_PHILOX, seed=12345)
    full = engine.hmc(x0, "iso_gauss", **kw)
    dr = full["draws"]
    assert dr.shape == (C, 1000, d) and np.isfinite(dr).all()
    # chains sharded over "GPUs": shard results are bit-identical to the single-call results
    for lo, hi in ((0, 512), (3584, 4096)):
        part = engine.hmc(x0[lo:hi], "iso_gauss", chain_offset=lo, **kw)
        assert np.array_equal(part["draws"], dr[lo:hi])
        assert np.array_equal(part["n_accept"], full["n_accept"][lo:hi])
    acc_rate = full["n_accept"].mean() / 1000
    assert acc_rate > 0.97, acc_rate
    last = dr[:, 500:, :]
    m = last.mean(axis=(0, 1))
    v = last.var(axis=(0, 1))
    # 4096*500 correlated draws per coordinate: |mean| and |var-1| well inside these bounds
    assert np.abs(m).max() < 0.02, np.abs(m).max()
    assert np.abs(v - 1).max() < 0.03, np.abs(v - 1).max()
    # chains must be distinct (different Philox substreams)
    assert np.abs(dr[0, -1] - dr[1, -1]).max() > 1e-3


@pytest.mark.parametrize("d", [514, 600, 1024, 1500, 2048])
def test_wide_hmc_four_warps_per_chain(engine, oracle, d):
    """512 < n_dim <= 2048 (BASELINE config 5's dimension sweep): one CTA of four warps per chain (hmc_wide.cu).
    The chain-wide reduction order differs from the oracle's, so STRICT is held to the contract tolerance."""
    C = 3
    rng = np.random.default_rng(d)
    x0 = rng.normal(size=(C, d))
    eps = 0.7 / d ** 0.25
    st = ol.Settings(n_burnin=3, n_keep=15, n_leap_steps=6, step_size=eps)
    for tname, tid, td in (("iso_gauss", ol.TGT_ISO_GAUSS, None), ("diag_gauss", ol.TGT_DIAG_GAUSS, np.linspace(0.5, 1.5, d))):
        od, oa, olp = _oracle_chains(oracle, ol.HMC, tid, td, x0, st, 55, ol.RNG_MT, ol.SUM_WARP)
        for a in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
            r = engine.hmc(x0, tname, target_data=td, n_leap_steps=6, step_size=eps, n_burnin=3, n_keep=15,
                           rng_mode=engine.api.RNG_MT19937_TAPE, seed=55, arith=a, want_logp=True)
            assert np.abs(r["draws"] - od).max() <= TOL, (tname, a)
            assert np.abs(r["logp"] - olp).max() <= 1e-9 * np.abs(olp).max()
            assert np.array_equal(r["n_accept"], oa)
        od, oa, _ = _oracle_chains(oracle, ol.HMC, tid, td, x0, st, 56, ol.RNG_PHILOX, ol.SUM_WARP, chain_offset=9)
        r = engine.hmc(x0, tname, target_data=td, n_leap_steps=6, step_size=eps, n_burnin=3, n_keep=15,
                       rng_mode=engine.api.RNG_PHILOX, seed=56, chain_offset=9)
        assert np.abs(r["draws"] - od).max() <= TOL and np.array_equal(r["n_accept"], oa)
    assert 0 < oa.max()
    # dense targets are not separable across warps: beyond 512 elements they run chain-batched (hmc_batched.cu, FAST arithmetic,
    # even n_dim); STRICT has no such path and is refused
    with pytest.raises(engine.McmcB200Error):
        engine.hmc(np.zeros((2, d)), "dense_gauss", target_data=np.eye(d), n_burnin=1, n_keep=1, arith=engine.api.ARITH_STRICT)
    r = engine.hmc(x0, "dense_gauss", target_data=np.eye(d), n_leap_steps=6, step_size=eps, n_burnin=3, n_keep=15, rng_mode=engine.api.RNG_PHILOX,
                   seed=56, chain_offset=9)
    od, oa, _ = _oracle_chains(oracle, ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, 56, ol.RNG_PHILOX, ol.SUM_WARP, chain_offset=9)
    assert np.abs(r["draws"] - od).max() <= TOL and np.array_equal(r["n_accept"], oa)   # precision matrix I == the iso-Gaussian


@pytest.mark.parametrize("tname", ["iso_gauss", "diag_gauss"])
def test_two_warps_per_chain_kernel_reproduces_the_production_kernel(engine, oracle, monkeypatch, tname):
    """Few chains (strong-scaling shards): hmc_duo.cu runs a chain on TWO warps — one generates draw t + 1's Philox / Box-Muller
    variates while the other runs draw t's trajectory (src/hmc.cpp:155-205) — coupled through shared memory and one named
    barrier per draw.  Same variates, same arithmetic: bit-identical to the one-warp-per-chain production kernel, and within
    1e-10 of the oracle; MCMCB200_HMC_DUO=0/1 forces the choice (default: n_leap != 10 and <= 2048 chains, where it is faster)."""
    rng = np.random.default_rng(8)
    for d, L, C in ((128, 10, 300), (64, 7, 37), (256, 3, 50), (128, 0, 9)):
        w = np.exp(rng.uniform(-0.5, 0.5, size=d)) if tname == "diag_gauss" else None
        x0 = rng.normal(size=(C, d))
        kw = dict(target_data=w, n_leap_steps=L, step_size=0.11, n_burnin=7, n_keep=25, rng_mode=engine.api.RNG_PHILOX, seed=31, chain_offset=3,
                  want_logp=True)
        monkeypatch.setenv("MCMCB200_HMC_DUO", "0")
        a = engine.hmc(x0, tname, **kw)
        monkeypatch.setenv("MCMCB200_HMC_DUO", "1")
        b = engine.hmc(x0, tname, **kw)
        monkeypatch.delenv("MCMCB200_HMC_DUO")
        c = engine.hmc(x0, tname, **kw)   # default choice
        assert np.array_equal(a["draws"], b["draws"]) and np.array_equal(a["n_accept"], b["n_accept"]) and np.array_equal(a["logp"], b["logp"]), (d, L)
        assert np.array_equal(b["draws"], c["draws"])
        st = ol.Settings(n_burnin=7, n_keep=25, n_leap_steps=L, step_size=0.11)
        tid = ol.TGT_DIAG_GAUSS if tname == "diag_gauss" else ol.TGT_ISO_GAUSS
        for ch in (0, C - 1):
            o = oracle.run_chain(ol.HMC, tid, w, x0[ch], st, seed=31, rng_mode=ol.RNG_PHILOX, chain_id=3 + ch, sum_mode=ol.SUM_WARP)
            assert np.abs(b["draws"][ch] - o["draws"]).max() <= TOL and b["n_accept"][ch] == o["n_accept"]


@pytest.mark.parametrize("tname", ["iso_gauss", "diag_gauss"])
def test_two_chains_per_warp_kernel_for_small_n_dim(engine, oracle, monkeypatch, tname):
    """n_dim <= 32 (BASELINE config 5's sweep point d = 32): hmc_half.cu puts two chains on one warp, 16 lanes each — same
    per-lane arithmetic, Philox counters and spare-bit uniform per half, 16-lane reductions.  Bit-identical to the
    warp-per-chain kernel (whose idle upper lanes only ever contributed zeros), odd chain counts and ragged n_dim included,
    and within 1e-10 of the oracle."""
    rng = np.random.default_rng(18)
    for d, L, C in ((32, 10, 301), (31, 4, 64), (17, 7, 9), (2, 3, 5), (1, 2, 4), (8, 0, 6)):
        w = np.exp(rng.uniform(-0.5, 0.5, size=d)) if tname == "diag_gauss" else None
        x0 = rng.normal(size=(C, d))
        kw = dict(target_data=w, n_leap_steps=L, step_size=0.13, n_burnin=5, n_keep=30, rng_mode=engine.api.RNG_PHILOX, seed=77, chain_offset=2,
                  want_logp=True)
        monkeypatch.setenv("MCMCB200_HMC_HALF", "0")
        a = engine.hmc(x0, tname, **kw)
        monkeypatch.setenv("MCMCB200_HMC_HALF", "1")
        b = engine.hmc(x0, tname, **kw)
        monkeypatch.delenv("MCMCB200_HMC_HALF")
        assert np.array_equal(a["draws"], b["draws"]), (d, L, C, np.abs(a["draws"] - b["draws"]).max())
        assert np.array_equal(a["n_accept"], b["n_accept"]) and np.array_equal(a["logp"], b["logp"])
        st = ol.Settings(n_burnin=5, n_keep=30, n_leap_steps=L, step_size=0.13)
        tid = ol.TGT_DIAG_GAUSS if tname == "diag_gauss" else ol.TGT_ISO_GAUSS
        for ch in (0, C - 1):
            o = oracle.run_chain(ol.HMC, tid, w, x0[ch], st, seed=77, rng_mode=ol.RNG_PHILOX, chain_id=2 + ch, sum_mode=ol.SUM_WARP)
            assert np.abs(b["draws"][ch] - o["draws"]).max() <= TOL and b["n_accept"][ch] == o["n_accept"]
