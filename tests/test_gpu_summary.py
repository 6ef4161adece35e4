"""On-device summaries of draws_out (include/mcmc_b200_summary.h) against numpy on the same draws."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _numpy_summary(a):
    C, T, d = a.shape
    cm = a.mean(axis=1)
    cv = a.var(axis=1, ddof=1) if T > 1 else np.full((C, d), np.nan)
    mean = cm.mean(axis=0)
    var = a.reshape(C * T, d).var(axis=0, ddof=1)
    W = cv.mean(axis=0)
    Bn = cm.var(axis=0, ddof=1) if C > 1 else np.full(d, np.nan)
    rhat = np.sqrt(((T - 1) / T * W + Bn) / W)
    return mean, var, rhat, cm, cv


@pytest.mark.parametrize("C,T,d", [(7, 33, 5), (16, 100, 64), (33, 57, 129), (12, 40, 300), (3, 2, 1)])
def test_summaries_match_numpy(engine, C, T, d):
    rng = np.random.default_rng(C * 1000 + d)
    a = rng.normal(size=(C, T, d)) * rng.uniform(0.1, 3.0, size=d) + 1e6 * rng.normal(size=d) + rng.normal(size=(C, 1, d)) * 0.3
    r = engine.api.summarize(a, per_chain=True)
    mean, var, rhat, cm, cv = _numpy_summary(a)
    assert np.allclose(r["chain_mean"], cm, rtol=1e-13, atol=0)
    assert np.allclose(r["chain_var"], cv, rtol=1e-9, atol=0)     # numpy's own two-pass variance at a 1e6 offset is good to ~1e-10
    assert np.allclose(r["mean"], mean, rtol=1e-13, atol=0)
    assert np.allclose(r["var"], var, rtol=1e-9, atol=0)
    assert np.allclose(r["rhat"], rhat, rtol=1e-8, atol=0)


def test_summary_of_a_run_stays_on_device(engine):
    """HMC with device-resident draws_out, summarised where it lies: moments of the target and R-hat ~ 1."""
    torch = pytest.importorskip("torch")
    C, d, nk = 512, 128, 200
    x0 = torch.from_numpy(ol.c2_initial(C, d)).cuda()
    draws = torch.empty((C, nk, d), dtype=torch.float64, device="cuda")
    r = engine.hmc(None, "iso_gauss", n_leap_steps=10, step_size=0.1, n_burnin=100, n_keep=nk, rng_mode=engine.api.RNG_PHILOX, seed=3,
                   initial_dev_ptr=x0.data_ptr(), n_chains=C, n_dim=d, draws_dev_ptr=draws.data_ptr(),
                   stream=torch.cuda.current_stream().cuda_stream)
    assert r["kernel_ms"] > 0
    s = engine.api.summarize(draws_dev_ptr=draws.data_ptr(), n_chains=C, n_keep=nk, n_dim=d, stream=torch.cuda.current_stream().cuda_stream)
    host = draws.cpu().numpy()
    mean, var, rhat, _, _ = _numpy_summary(host)
    assert np.allclose(s["mean"], mean, rtol=0, atol=1e-13) and np.allclose(s["var"], var, rtol=1e-12) and np.allclose(s["rhat"], rhat, rtol=1e-10)
    assert np.abs(s["mean"]).max() < 0.05 and np.abs(s["var"] - 1).max() < 0.05 and np.abs(s["rhat"] - 1).max() < 0.05


def test_summary_argument_errors(engine):
    with pytest.raises(engine.McmcB200Error):
        engine.api.summarize(np.zeros((2, 0, 3)))
