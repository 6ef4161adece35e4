"""GPU tests for the chain-batched MALA path (mala_wide.cu): one fp64 tensor-core GEMM per draw for all chains.
BASELINE config 3 shape: Bayesian linear regression posterior on sufficient statistics, M = I."""
import numpy as np
import pytest

import oracle_lib as ol
from test_gpu_hmc import TOL

pytestmark = pytest.mark.gpu


def linreg_problem(d, n_obs, seed=7, sigma=0.5, tau=10.0):
    """A = X'X/sigma^2 + I/tau^2, b = X'y/sigma^2, y = X beta* + sigma N(0,1), beta*_j = sin j."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n_obs, d))
    beta = np.sin(np.arange(d, dtype=np.float64))
    y = X @ beta + sigma * rng.normal(size=n_obs)
    A = X.T @ X / sigma ** 2 + np.eye(d) / tau ** 2
    A = (A + A.T) / 2
    b = X.T @ y / sigma ** 2
    lam = np.linalg.eigvalsh(A)[-1]
    return A, b, 0.5 / np.sqrt(lam), beta


def _oracle(oracle, tid, td, x0s, st, seed, rng_mode, chain_offset=0):
    out, acc = [], []
    for c in range(x0s.shape[0]):
        s = seed + chain_offset + c if rng_mode == ol.RNG_MT else seed
        o = oracle.run_chain(ol.MALA, tid, td, x0s[c], st, seed=s, rng_mode=rng_mode, chain_id=chain_offset + c, sum_mode=ol.SUM_SEQ)
        out.append(o["draws"]); acc.append(o["n_accept"])
    return np.stack(out), np.array(acc)


@pytest.mark.parametrize("d,C", [(96, 300), (130, 257), (512, 256), (1024, 300)])
def test_linreg_vs_oracle(engine, oracle, d, C):
    """>= 256 chains (or d > 512) selects the chain-batched path; first 6 chains are checked against the CPU oracle."""
    A, b, eps, _ = linreg_problem(d, 2 * d)
    td = np.concatenate([A.ravel(), b])
    rng = np.random.default_rng(d)
    x0 = rng.normal(size=(C, d)) * 0.1
    n_keep = 12 if d >= 512 else 30
    st = ol.Settings(n_burnin=3, n_keep=n_keep, step_size=eps)
    for rng_mode, orng in ((engine.api.RNG_MT19937_TAPE, ol.RNG_MT), (engine.api.RNG_PHILOX, ol.RNG_PHILOX)):
        r = engine.mala(x0, "linreg", target_data=td, step_size=eps, n_burnin=3, n_keep=n_keep, rng_mode=rng_mode, seed=77,
                        chain_offset=5)
        assert r["kernel_launches"] == 2 * (3 + n_keep) + 2  # gemm + rows per draw: the batched path really ran
        od, oa = _oracle(oracle, ol.TGT_LINREG, td, x0[:6], st, 77, orng, chain_offset=5)
        scale = max(1.0, np.abs(od).max())
        assert np.abs(r["draws"][:6] - od).max() <= TOL * scale, np.abs(r["draws"][:6] - od).max()
        assert np.array_equal(r["n_accept"][:6], oa)
        assert oa.max() > 0


def test_dense_gauss_and_path_selection(engine, oracle):
    rng = np.random.default_rng(3)
    d = 64
    a = rng.normal(size=(d, d)); P = a @ a.T / d + np.eye(d); P = (P + P.T) / 2
    st = ol.Settings(n_burnin=2, n_keep=20, step_size=0.3)
    x0 = rng.normal(size=(256, d))
    r = engine.mala(x0, "dense_gauss", target_data=P, step_size=0.3, n_burnin=2, n_keep=20, rng_mode=engine.api.RNG_MT19937_TAPE, seed=9)
    assert r["kernel_launches"] > 1
    od, oa = _oracle(oracle, ol.TGT_DENSE_GAUSS, P.ravel(), x0[:4], st, 9, ol.RNG_MT)
    assert np.abs(r["draws"][:4] - od).max() <= TOL and np.array_equal(r["n_accept"][:4], oa)
    # few chains -> register-resident kernel (one launch); same answer
    r2 = engine.mala(x0[:4], "dense_gauss", target_data=P, step_size=0.3, n_burnin=2, n_keep=20, rng_mode=engine.api.RNG_MT19937_TAPE,
                     seed=9)
    assert r2["kernel_launches"] == 1
    assert np.abs(r2["draws"] - od).max() <= TOL
    # a dense preconditioner at d > 512 is not built: refuse
    with pytest.raises(engine.McmcB200Error) as e:
        engine.mala(np.zeros((4, 600)), "dense_gauss", target_data=np.eye(600), precond_mat=np.eye(600), n_burnin=1, n_keep=1)
    assert e.value.code == engine.api.ERR_UNSUPPORTED


def test_c3_full_size_properties(engine):
    """BASELINE config 3 at full size: d=1024, 16384 chains, 20 + 100 draws.  Checked through size-independent
    properties: finite draws, a sane acceptance rate, the ensemble mean drifting to the posterior mean A^-1 b, and
    sharding invariance of a block of chains."""
    d, C = 1024, 16384
    A, b, eps, beta = linreg_problem(d, 4096)
    td = np.concatenate([A.ravel(), b])
    post_mean = np.linalg.solve(A, b)
    rng = np.random.default_rng(0)
    x0 = post_mean[None, :] + 0.05 * rng.normal(size=(C, d))
    r = engine.mala(x0, "linreg", target_data=td, step_size=eps, n_burnin=20, n_keep=100, rng_mode=engine.api.RNG_PHILOX, seed=2024)
    dr = r["draws"]
    assert dr.shape == (C, 100, d) and np.isfinite(dr[:, -1]).all()
    acc = r["n_accept"].mean() / 100
    assert 0.3 < acc < 0.999, acc
    # started around the posterior mean with a too-wide spread: the ensemble mean stays at A^-1 b
    err = np.abs(dr[:, -1].mean(axis=0) - post_mean).max()
    assert err < 0.01, err
    part = engine.mala(x0[4096:4352], "linreg", target_data=td, step_size=eps, n_burnin=20, n_keep=100, rng_mode=engine.api.RNG_PHILOX,
                       seed=2024, chain_offset=4096)
    assert np.abs(part["draws"] - dr[4096:4352]).max() <= TOL  # same chains, different GEMM tile positions
    print("C3 kernel time %.1f ms for %d chains x 120 draws" % (r["kernel_ms"], C))
