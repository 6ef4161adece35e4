"""GPU parity tests for mcmc::de (src/de.cpp:30-271) through the C ABI (mcmcb200_de_run, csrc/de.cu).

Gate: the restated sampler oracle_run_de, bit-identical to the unmodified src/de.cpp (tests/test_oracle_vs_reference.py).
  * MCMCB200_RNG_MT19937_TAPE: the library replays the reference's own stream on the host (host_de_tape); STRICT
    arithmetic must reproduce the reference's draws — bit for bit where no transcendental is involved in the target
    (log z of the accept test is compared against log-density differences; a decision could only flip at a rounding-level
    margin, never observed) — and the accept counts; FAST within 1e-10.
  * MCMCB200_RNG_PHILOX: the device's own counter-based stream; the test rebuilds the population's tape from the raw device
    stream (mcmcb200_philox_stream, itself checked against the oracle elsewhere) and the oracle replays it.
Many populations per call must equal the same populations run one by one (global population ids)."""
import numpy as np
import pytest

import golden_util
import oracle_lib as ol
from test_gpu_hmc import TOL
from test_oracle_vs_reference import _de_cases

pytestmark = pytest.mark.gpu
NAMES = {ol.TGT_ISO_GAUSS: "iso_gauss", ol.TGT_DIAG_GAUSS: "diag_gauss", ol.TGT_DENSE_GAUSS: "dense_gauss", ol.TGT_FUNNEL: "funnel",
         ol.TGT_NORMAL_MODEL: "normal_model", ol.TGT_LINREG: "linreg"}
BITEXACT = {ol.TGT_ISO_GAUSS, ol.TGT_DIAG_GAUSS, ol.TGT_DENSE_GAUSS}


def _kw(st):
    return dict(n_pop=st["n_pop"], jumps=st["jumps"], par_b=st["par_b"], par_gamma_jump=st["par_gamma_jump"], initial_lb=st["initial_lb"],
                initial_ub=st["initial_ub"], n_burnin=st["n_burnin"], n_keep=st["n_keep"], lower_bounds=st["lower_bounds"], upper_bounds=st["upper_bounds"])


def test_reference_stream_vs_reference_and_oracle(engine, oracle, reference):
    for name, tid, tdata, x0, st, seed in _de_cases():
        ref, acc = reference.run_de(tid, tdata, x0, st, seed)
        for arith in (engine.api.ARITH_STRICT, engine.api.ARITH_FAST):
            r = engine.de(np.asarray(x0, dtype=np.float64)[None, :], NAMES[tid], target_data=tdata, rng_mode=engine.api.RNG_MT19937_TAPE, seed=seed,
                          arith=arith, **_kw(st))
            assert r["draws"].shape == (1,) + ref.shape
            assert np.abs(r["draws"][0] - ref).max() <= TOL, (name, arith, np.abs(r["draws"][0] - ref).max())
            assert r["n_accept"][0] == acc, (name, arith)
            if arith == engine.api.ARITH_STRICT and tid in BITEXACT and st["lower_bounds"] is None:
                assert np.array_equal(r["draws"][0], ref), name


def test_golden_cases(engine):
    g = golden_util.load()
    for c in g["de_cases"]:
        st = ol.DeSettings(**c["st"])
        if "lower" in c:
            st["lower_bounds"] = np.array([float.fromhex(h) for h in c["lower"]])
            st["upper_bounds"] = np.array([float.fromhex(h) for h in c["upper"]])
        want = np.array([float.fromhex(h) for h in c["draws_hex"]]).reshape(c["draws_shape"])
        r = engine.de(np.asarray(c["x0"], dtype=np.float64)[None, :], NAMES[c["target"]], target_data=c["tdata"], rng_mode=engine.api.RNG_MT19937_TAPE,
                      seed=c["seed"], arith=engine.api.ARITH_STRICT, **_kw(st))
        assert np.abs(r["draws"][0] - want).max() <= TOL, c["name"]
        assert r["n_accept"][0] == c["n_accept"], c["name"]


def _philox_tape(engine, seed, pop, n_pop, d, n_total, b):
    """The population's variates in the tape layout, from the raw device Philox stream (uniform #k of word g*n_pop+i)."""
    out = []
    u0 = engine.api.philox_stream(seed, pop, -1, 2, n_pop * d + 1)[2:]
    out.extend(u0[1:1 + n_pop * d])
    for g in range(n_total):
        for i in range(n_pop):
            u = engine.api.philox_stream(seed, pop, g * n_pop + i, 2, d + 4)[2:]
            c1 = min(int(u[1] * (n_pop - 1)), n_pop - 2)
            c1 += 1 if c1 >= i else 0
            c2 = min(int(u[2] * (n_pop - 2)), n_pop - 3)
            s0, s1 = min(i, c1), max(i, c1)
            c2 += 1 if c2 >= s0 else 0
            c2 += 1 if c2 >= s1 else 0
            out += [float(c1), float(c2)] + list((2.0 * b) * u[3:3 + d] - b) + [u[d + 3]]
    return np.array(out)


@pytest.mark.parametrize("case", [0, 1, 3])
def test_philox_stream_vs_oracle(engine, oracle, case):
    name, tid, tdata, x0, st, seed = _de_cases()[case]
    d, n_pop, n_total = len(x0), st["n_pop"], st["n_burnin"] + st["n_keep"]
    pop_id = 5
    tape = _philox_tape(engine, 4242, pop_id, n_pop, d, n_total, st["par_b"])
    o = oracle.run_de(tid, tdata, x0, st, rng_mode=ol.RNG_TAPE, tape=tape, sum_mode=ol.SUM_WARP)
    for arith, tol in ((engine.api.ARITH_STRICT, TOL), (engine.api.ARITH_FAST, TOL)):
        r = engine.de(np.asarray(x0, dtype=np.float64)[None, :], NAMES[tid], target_data=tdata, rng_mode=engine.api.RNG_PHILOX, seed=4242,
                      chain_offset=pop_id, arith=arith, **_kw(st))
        assert np.abs(r["draws"][0] - o["draws"]).max() <= tol, (name, arith)
        assert r["n_accept"][0] == o["n_accept"], (name, arith)
    # the same tape handed to the kernel as a USER_TAPE
    r = engine.de(np.asarray(x0, dtype=np.float64)[None, :], NAMES[tid], target_data=tdata, rng_mode=engine.api.RNG_USER_TAPE, tape=tape[None, :],
                  arith=engine.api.ARITH_STRICT, **_kw(st))
    assert np.abs(r["draws"][0] - o["draws"]).max() <= TOL and r["n_accept"][0] == o["n_accept"]
    with pytest.raises(engine.McmcB200Error):   # a tape that is too short is refused, not read past its end
        engine.de(np.asarray(x0, dtype=np.float64)[None, :], NAMES[tid], target_data=tdata, rng_mode=engine.api.RNG_USER_TAPE, tape=tape[None, :-3],
                  arith=engine.api.ARITH_STRICT, **_kw(st))


def test_many_populations_equal_single_runs_and_sample_the_target(engine):
    """64 populations of 24 members on a d=16 diagonal Gaussian in one call: population p equals a single-population call with
    chain_offset=p (global ids), the ensemble reproduces the target's variances, sizes beyond the kernel set are refused."""
    d, P, n_pop = 16, 64, 24
    w = np.linspace(0.5, 3.0, d)
    x0 = np.zeros((P, d))
    kw = dict(target_data=w, n_pop=n_pop, n_burnin=400, n_keep=200, par_b=1e-4, rng_mode=engine.api.RNG_PHILOX, seed=9)
    r = engine.de(x0, "diag_gauss", **kw)
    assert r["draws"].shape == (P, 200, n_pop, d)
    for p in (0, 17, 63):
        one = engine.de(x0[p:p + 1], "diag_gauss", chain_offset=p, **kw)
        assert np.array_equal(one["draws"][0], r["draws"][p]) and one["n_accept"][0] == r["n_accept"][p]
    var = r["draws"].reshape(-1, d).var(axis=0)
    assert np.abs(var * w - 1.0).max() < 0.15, var * w
    acc = r["n_accept"].mean() / (200 * n_pop)
    assert 0.15 < acc < 0.7, acc
    with pytest.raises(engine.McmcB200Error):
        engine.de(np.zeros((1, 4)), "iso_gauss", n_pop=2, n_burnin=1, n_keep=1)
    with pytest.raises(engine.McmcB200Error):
        engine.de(np.zeros((1, 600)), "iso_gauss", n_pop=8, n_burnin=1, n_keep=1)
