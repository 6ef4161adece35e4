"""GPU parity tests for the chain-batched HMC path (csrc/hmc_batched.cu): dense mass matrix and / or dense quadratic target,
every d x d product of the trajectory as one fp64 tensor-core (DMMA) GEMM over all chains (src/hmc.cpp:57-59,158-171).
FAST arithmetic: held to the 1e-10 contract against the oracle (warp summation order) and against the warp-per-chain kernel."""
import numpy as np
import pytest

import oracle_lib as ol
from test_gpu_hmc import _sym_pd, TOL

pytestmark = pytest.mark.gpu


def _targets(rng, d):
    a = rng.normal(size=(d, d))
    A = a @ a.T / d + np.eye(d)
    A = (A + A.T) / 2
    return {"iso_gauss": (ol.TGT_ISO_GAUSS, None), "diag_gauss": (ol.TGT_DIAG_GAUSS, np.exp(rng.uniform(-0.5, 0.5, size=d))),
            "dense_gauss": (ol.TGT_DENSE_GAUSS, A.ravel()), "linreg": (ol.TGT_LINREG, np.concatenate([A.ravel(), rng.normal(size=d)]))}


@pytest.mark.parametrize("tname,dense_mass", [("iso_gauss", True), ("diag_gauss", True), ("dense_gauss", False), ("dense_gauss", True), ("linreg", True)])
def test_batched_vs_oracle_and_warp_kernel(engine, oracle, monkeypatch, tname, dense_mass):
    import torch

    rng = np.random.default_rng(len(tname) + dense_mass)
    d, C, L, eps = 40, 600, 6, 0.12
    tid, td = _targets(rng, d)[tname]
    M = _sym_pd(rng, d, 0.5) if dense_mass else None
    x0 = rng.normal(size=(C, d)) * 0.5
    st = ol.Settings(n_burnin=3, n_keep=20, n_leap_steps=L, step_size=eps, precond=M)
    kw = dict(target_data=td, n_leap_steps=L, step_size=eps, precond_mat=M, n_burnin=3, n_keep=20, want_logp=True)
    stream = torch.cuda.Stream()
    for mode, seed in ((engine.api.RNG_PHILOX, 5), (engine.api.RNG_MT19937_TAPE, 6)):
        with torch.cuda.stream(stream):   # a real stream: the per-draw launch sequence is captured in a CUDA graph and replayed
            r = engine.hmc(x0, tname, rng_mode=mode, seed=seed, stream=stream.cuda_stream, **kw)
        assert r["kernel_launches"] > 23   # the chain-batched path (one launch would be the warp kernel)
        rd = engine.hmc(x0, tname, rng_mode=mode, seed=seed, **kw)   # legacy default stream: launches issued directly
        assert np.array_equal(r["draws"], rd["draws"]) and np.array_equal(r["n_accept"], rd["n_accept"])
        monkeypatch.setenv("MCMCB200_HMC_BATCHED", "0")
        w = engine.hmc(x0, tname, rng_mode=mode, seed=seed, **kw)
        monkeypatch.delenv("MCMCB200_HMC_BATCHED")
        assert w["kernel_launches"] == 1
        assert np.abs(r["draws"] - w["draws"]).max() <= TOL and np.array_equal(r["n_accept"], w["n_accept"])
        assert np.abs(r["logp"] - w["logp"]).max() <= 1e-9
        for c in (0, 311, C - 1):
            o = (oracle.run_chain(ol.HMC, tid, td, x0[c], st, seed=seed, rng_mode=ol.RNG_PHILOX, chain_id=c, sum_mode=ol.SUM_WARP) if mode == engine.api.RNG_PHILOX
                 else oracle.run_chain(ol.HMC, tid, td, x0[c], st, seed=seed + c, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_WARP))
            assert np.abs(r["draws"][c] - o["draws"]).max() <= TOL, (tname, mode, c)
            assert r["n_accept"][c] == o["n_accept"]


def test_beyond_the_register_kernels_and_edge_cases(engine, oracle):
    """Dense targets / dense mass above n_dim = 512 exist only on the chain-batched path; ragged chain counts, L = 0,
    STRICT and odd n_dim are handled or refused loudly."""
    rng = np.random.default_rng(9)
    d, C = 768, 70
    a = rng.normal(size=(d, d))
    A = a @ a.T / d + np.eye(d)
    A = (A + A.T) / 2
    M = _sym_pd(rng, d, 0.5)
    x0 = rng.normal(size=(C, d)) * 0.3
    st = ol.Settings(n_burnin=1, n_keep=5, n_leap_steps=4, step_size=0.05, precond=M)
    r = engine.hmc(x0, "dense_gauss", target_data=A.ravel(), n_leap_steps=4, step_size=0.05, precond_mat=M, n_burnin=1, n_keep=5,
                   rng_mode=engine.api.RNG_PHILOX, seed=3, chain_offset=100)
    for c in (0, C - 1):
        o = oracle.run_chain(ol.HMC, ol.TGT_DENSE_GAUSS, A.ravel(), x0[c], st, seed=3, rng_mode=ol.RNG_PHILOX, chain_id=100 + c, sum_mode=ol.SUM_WARP)
        assert np.abs(r["draws"][c] - o["draws"]).max() <= TOL and r["n_accept"][c] == o["n_accept"]
    # L = 0: the proposal is the current state, every draw is "accepted" (src/hmc.cpp:164-191 with an empty loop)
    z = engine.hmc(x0[:, :64].copy(), "dense_gauss", target_data=np.eye(64).ravel(), n_leap_steps=0, step_size=0.1, n_burnin=0, n_keep=3,
                   rng_mode=engine.api.RNG_PHILOX, seed=1, n_chains=None)
    assert np.array_equal(z["draws"][:, 0], x0[:, :64]) and (z["n_accept"] == 3).all()
    with pytest.raises(engine.McmcB200Error) as e:   # STRICT has no GEMM order: beyond the warp kernels it is refused
        engine.hmc(x0, "dense_gauss", target_data=A.ravel(), n_burnin=1, n_keep=1, arith=engine.api.ARITH_STRICT)
    assert e.value.code == engine.api.ERR_UNSUPPORTED
