"""GPU parity tests for the chain-batched NUTS path (csrc/nuts_batched.cu; src/nuts.cpp:30-332, nuts.ipp:30-241): all chains
advance in lock-step rounds, a round = one fp64 tensor-core (DMMA) GEMM for every chain's pending gradient product + one launch
of the resumable per-chain state machine.  FAST arithmetic: held to the 1e-10 contract (adaptation off) / ADAPT_TOL (dual
averaging on, see test_gpu_nuts.py) against the oracle's literal recursion and against the persistent warp-per-chain kernel."""
import numpy as np
import pytest

import oracle_lib as ol
from test_gpu_hmc import TOL
from test_gpu_nuts import ADAPT_TOL

pytestmark = pytest.mark.gpu


def _dense(rng, d, with_b):
    a = rng.normal(size=(d, d))
    A = a @ a.T / d + np.eye(d)
    A = (A + A.T) / 2
    if with_b:
        return ol.TGT_LINREG, "linreg", np.concatenate([A.ravel(), rng.normal(size=d)])
    return ol.TGT_DENSE_GAUSS, "dense_gauss", A.ravel()


@pytest.mark.parametrize("with_b", [False, True])
def test_batched_rounds_vs_persistent_kernel_and_oracle(engine, oracle, monkeypatch, with_b):
    rng = np.random.default_rng(3 + with_b)
    d, C = 48, 300
    tid, tname, td = _dense(rng, d, with_b)
    x0 = rng.normal(size=(C, d))
    for n_adapt, eps0, tol in ((0, 0.12, TOL), (30, 1.0, ADAPT_TOL)):
        kw = dict(target_data=td, n_burnin=30, n_keep=30, n_adapt_draws=n_adapt, step_size=eps0, rng_mode=engine.api.RNG_PHILOX, seed=91,
                  chain_offset=7, want_logp=True)
        monkeypatch.setenv("MCMCB200_NUTS_BATCHED", "1")
        r = engine.nuts(x0, tname, **kw)
        assert r["kernel_launches"] > 100   # rounds of (GEMM, step kernel), not one persistent launch
        monkeypatch.setenv("MCMCB200_NUTS_BATCHED", "0")
        monkeypatch.setenv("MCMCB200_NUTS_COOP", "0")
        w = engine.nuts(x0, tname, **kw)
        monkeypatch.delenv("MCMCB200_NUTS_BATCHED")
        monkeypatch.delenv("MCMCB200_NUTS_COOP")
        assert w["kernel_launches"] == 1
        linf = np.abs(r["draws"] - w["draws"]).max(axis=(1, 2))
        assert linf.max() <= tol, (n_adapt, linf.max(), int((linf > tol).sum()))
        assert np.array_equal(r["n_accept"], w["n_accept"])
        assert np.array_equal(r["n_leapfrog"], w["n_leapfrog"])
        assert np.allclose(r["step_size"], w["step_size"], rtol=1e3 * tol, atol=0)
        assert np.abs(r["logp"] - w["logp"]).max() <= 1e3 * tol
        st = ol.Settings(n_burnin=30, n_keep=30, n_adapt_draws=n_adapt, step_size=eps0)
        for c in (0, 151, C - 1):
            o = oracle.run_chain(ol.NUTS, tid, td, x0[c], st, seed=91, rng_mode=ol.RNG_PHILOX, chain_id=7 + c, sum_mode=ol.SUM_WARP)
            assert np.abs(r["draws"][c] - o["draws"]).max() <= 50 * tol, (n_adapt, c, np.abs(r["draws"][c] - o["draws"]).max())
            assert r["n_accept"][c] == o["n_accept"]


def test_batched_rounds_on_the_reference_stream_tape(engine, oracle, monkeypatch):
    """USER_TAPE: the oracle runs the literal recursion on std::mt19937_64 and records every variate; the coroutine replays it
    (the tape cursor survives the rounds in the chain's control block) and reports how much of the tape it consumed."""
    rng = np.random.default_rng(12)
    d, C = 20, 5
    tid, tname, td = _dense(rng, d, False)
    x0 = rng.normal(size=(C, d))
    st = ol.Settings(n_burnin=5, n_keep=25, n_adapt_draws=0, step_size=0.1, max_tree_depth=7)
    tapes, od, oa = [], [], []
    for c in range(C):
        o = oracle.run_chain(ol.NUTS, tid, td, x0[c], st, seed=40 + c, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_WARP, record_tape=2_000_000)
        tapes.append(o["tape"]); od.append(o["draws"]); oa.append(o["n_accept"])
    L = max(len(t) for t in tapes) + 8
    tape = np.zeros((C, L))
    for c in range(C):
        tape[c, :len(tapes[c])] = tapes[c]
    monkeypatch.setenv("MCMCB200_NUTS_BATCHED", "1")
    r = engine.nuts(x0, tname, target_data=td, n_burnin=5, n_keep=25, n_adapt_draws=0, step_size=0.1, max_tree_depth=7,
                    rng_mode=engine.api.RNG_USER_TAPE, tape=tape)
    assert r["kernel_launches"] > 100
    assert np.abs(r["draws"] - np.stack(od)).max() <= TOL
    assert np.array_equal(r["n_accept"], np.array(oa))
    # a tape that is too short is reported, not read past its end
    with pytest.raises(engine.McmcB200Error):
        engine.nuts(x0, tname, target_data=td, n_burnin=0, n_keep=5, n_adapt_draws=0, step_size=0.05, rng_mode=engine.api.RNG_USER_TAPE,
                    tape=np.full((C, 2 * d + 6), 0.3))
    monkeypatch.delenv("MCMCB200_NUTS_BATCHED")


def test_batched_is_the_default_for_many_chains_and_shards_consistently(engine):
    """>= 512 chains on a dense target take the batched path by default; a shard of the call (global chain ids through
    chain_offset) reproduces the same chains bit for bit, whatever the other chains of its call are doing."""
    rng = np.random.default_rng(5)
    d, C = 64, 1100
    q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    P = (q / np.logspace(0, 2, d)) @ q.T
    P = (P + P.T) / 2
    x0 = rng.normal(size=(C, d))
    kw = dict(target_data=P.ravel(), n_burnin=10, n_keep=10, n_adapt_draws=10, rng_mode=engine.api.RNG_PHILOX, seed=8)
    r = engine.nuts(x0, "dense_gauss", **kw)
    assert r["kernel_launches"] == 2   # n_dim = 64 is a full tile: the persistent kernel (tables + one launch)
    assert np.isfinite(r["draws"]).all() and (r["n_leapfrog"] > 0).all()
    import os
    os.environ["MCMCB200_NUTS_BATCHED"] = "1"
    try:
        part = engine.nuts(x0[1000:], "dense_gauss", chain_offset=1000, **kw)
    finally:
        del os.environ["MCMCB200_NUTS_BATCHED"]
    assert np.array_equal(part["draws"], r["draws"][1000:]) and np.array_equal(part["n_leapfrog"], r["n_leapfrog"][1000:])


def test_deep_trees_use_the_global_tables_and_full_tiles(engine, monkeypatch):
    """max_tree_depth = 12: the summary table (2299 entries) and the per-state leaf records no longer fit shared memory and live in
    the chain's global area (the MEMO_SH = false build of the step kernel); n_dim = 128 = a full register tile (no bounds
    predicates).  Tiny fixed step so that trees really get deep; the persistent kernel is the reference."""
    rng = np.random.default_rng(17)
    for d, depth, eps, C in ((128, 12, 0.004, 40), (256, 10, 0.05, 24), (10, 12, 0.002, 9)):
        q, _ = np.linalg.qr(rng.normal(size=(d, d)))
        P = (q / np.logspace(0, 1, d)) @ q.T
        P = (P + P.T) / 2
        x0 = rng.normal(size=(C, d))
        kw = dict(target_data=P.ravel(), n_burnin=0, n_keep=5, n_adapt_draws=0, step_size=eps, max_tree_depth=depth, rng_mode=engine.api.RNG_PHILOX, seed=3)
        monkeypatch.setenv("MCMCB200_NUTS_BATCHED", "1")
        r = engine.nuts(x0, "dense_gauss", **kw)
        monkeypatch.setenv("MCMCB200_NUTS_BATCHED", "0")
        monkeypatch.setenv("MCMCB200_NUTS_COOP", "0")
        w = engine.nuts(x0, "dense_gauss", **kw)
        monkeypatch.delenv("MCMCB200_NUTS_BATCHED")
        monkeypatch.delenv("MCMCB200_NUTS_COOP")
        assert (r["kernel_launches"] > 50 or (depth <= 10 and r["kernel_launches"] == 2)) and w["kernel_launches"] == 1
        assert np.array_equal(r["n_leapfrog"], w["n_leapfrog"]) and np.array_equal(r["n_accept"], w["n_accept"]), (d, depth)
        assert np.abs(r["draws"] - w["draws"]).max() <= TOL, (d, depth, np.abs(r["draws"] - w["draws"]).max())
        if depth == 12 and d == 128:
            assert r["n_leapfrog"].max() > 5 * 100   # deep trees did occur (a depth-9 doubling alone has 46 distinct states)


def test_persistent_variant_reproduces_the_rounds_bit_for_bit(engine, oracle, monkeypatch):
    """Full tiles (n_dim = 64, 128, 256) with max_tree_depth <= 10 run as ONE persistent kernel — 16 chains per CTA, the coroutine
    state and the pending vectors stay in shared memory, the CTA computes its own chains' products with the same DMMA
    accumulation order, free warps take the next chain from a global counter.  MCMCB200_NUTS_PERSIST=0 keeps the launched
    rounds.  Same operations in the same order: identical bits, whatever the chain-to-CTA assignment turns out to be; also
    with two independently running groups of 8 chains per CTA (MCMCB200_NUTS_PERSIST_NH=2), and on the reference's stream."""
    rng = np.random.default_rng(29)
    for d, C, n_adapt, eps0 in ((64, 700, 20, 1.0), (256, 90, 0, 0.08), (128, 33, 10, 1.0)):
        q, _ = np.linalg.qr(rng.normal(size=(d, d)))
        P = (q / np.logspace(0, 2, d)) @ q.T
        P = (P + P.T) / 2
        x0 = rng.normal(size=(C, d))
        kw = dict(target_data=P.ravel(), n_burnin=20, n_keep=20, n_adapt_draws=n_adapt, step_size=eps0, rng_mode=engine.api.RNG_PHILOX, seed=4,
                  want_logp=True)
        monkeypatch.setenv("MCMCB200_NUTS_BATCHED", "1")
        monkeypatch.setenv("MCMCB200_NUTS_PERSIST", "0")
        r = engine.nuts(x0, "dense_gauss", **kw)
        monkeypatch.delenv("MCMCB200_NUTS_PERSIST")
        # chains per CTA (8: the default below 8 x #SMs chains, 16: above) x independently running sub-groups per CTA
        for nw, nh in (("8", "1"), ("16", "1"), ("16", "2")):
            monkeypatch.setenv("MCMCB200_NUTS_PERSIST_NW", nw)
            monkeypatch.setenv("MCMCB200_NUTS_PERSIST_NH", nh)
            p = engine.nuts(x0, "dense_gauss", **kw)
            assert r["kernel_launches"] > 100 and p["kernel_launches"] == 2
            assert np.array_equal(r["draws"], p["draws"]), (d, nw, nh, np.abs(r["draws"] - p["draws"]).max())
            assert np.array_equal(r["n_accept"], p["n_accept"]) and np.array_equal(r["n_leapfrog"], p["n_leapfrog"])
            assert np.array_equal(r["step_size"], p["step_size"]) and np.array_equal(r["logp"], p["logp"])
        monkeypatch.delenv("MCMCB200_NUTS_PERSIST_NH")
        monkeypatch.delenv("MCMCB200_NUTS_PERSIST_NW")
        monkeypatch.delenv("MCMCB200_NUTS_BATCHED")
    # the reference's own stream (oracle-recorded tape) through the persistent kernel, linreg target
    d, C = 64, 3
    tid, tname, td = _dense(rng, d, True)
    x0 = rng.normal(size=(C, d))
    st = ol.Settings(n_burnin=3, n_keep=12, n_adapt_draws=0, step_size=0.08, max_tree_depth=8)
    tapes, od, oa = [], [], []
    for c in range(C):
        o = oracle.run_chain(ol.NUTS, tid, td, x0[c], st, seed=60 + c, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_WARP, record_tape=2_000_000)
        tapes.append(o["tape"]); od.append(o["draws"]); oa.append(o["n_accept"])
    tape = np.zeros((C, max(len(t) for t in tapes) + 8))
    for c in range(C):
        tape[c, :len(tapes[c])] = tapes[c]
    monkeypatch.setenv("MCMCB200_NUTS_BATCHED", "1")
    r = engine.nuts(x0, tname, target_data=td, n_burnin=3, n_keep=12, n_adapt_draws=0, step_size=0.08, max_tree_depth=8,
                    rng_mode=engine.api.RNG_USER_TAPE, tape=tape)
    monkeypatch.delenv("MCMCB200_NUTS_BATCHED")
    assert r["kernel_launches"] == 2
    assert np.abs(r["draws"] - np.stack(od)).max() <= TOL and np.array_equal(r["n_accept"], np.array(oa))
