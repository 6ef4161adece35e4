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
