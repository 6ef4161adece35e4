"""Randomised parity sweep: seeded random configurations (dimension — odd, ragged and full tiles —, chain count, target,
step size, trajectory length, dense / identity mass or proposal covariance, RNG mode, chain offset) of HMC, MALA and RWMH
through the C ABI against the CPU oracle.  STRICT arithmetic on the reference's random stream must be bit-exact for targets
without transcendentals, FAST arithmetic / Philox within the contract tolerance (1e-10), accept counts identical."""
import numpy as np
import pytest

import oracle_lib as ol
from test_gpu_bounds import TNAME
from test_gpu_hmc import _oracle_chains, _sym_pd, TOL

pytestmark = pytest.mark.gpu

DIMS = [1, 2, 3, 7, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256, 257, 300, 511, 512]


def _target(rng, d):
    kind = rng.integers(0, 4)
    if kind == 0:
        return ol.TGT_ISO_GAUSS, None
    if kind == 1:
        return ol.TGT_DIAG_GAUSS, np.exp(rng.uniform(-1.0, 1.0, d))
    dd = min(d, 96)   # dense targets cost O(d^2) per gradient in the oracle: keep them small
    if kind == 2:
        return ol.TGT_DENSE_GAUSS, _sym_pd(rng, dd, 1.0).ravel()
    return ol.TGT_LINREG, np.concatenate([_sym_pd(rng, dd, 2.0).ravel(), rng.normal(size=dd)])


def _configs(sampler, n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        d = int(rng.choice(DIMS))
        tid, td = _target(rng, d)
        if tid in (ol.TGT_DENSE_GAUSS, ol.TGT_LINREG):
            d = min(d, 96)
        C = int(rng.integers(1, 10))
        dense = d <= 48 and rng.random() < 0.3
        mat = _sym_pd(rng, d, 0.6) if dense else None
        scale = float(rng.uniform(0.3, 1.2)) / d ** (0.25 if sampler == ol.HMC else 0.5 if sampler == ol.RWMH else 1 / 3)
        st = dict(n_burnin=int(rng.integers(0, 6)), n_keep=int(rng.integers(1, 25)), step_size=scale, precond=mat)
        if sampler == ol.HMC:
            st["n_leap_steps"] = int(rng.choice([0, 1, 2, 5, 10, 13]))
        out.append(dict(d=d, tid=tid, td=td, C=C, st=ol.Settings(**st), x0=rng.normal(size=(C, d)) * 0.7, seed=int(rng.integers(1, 2 ** 40)),
                        offset=int(rng.integers(0, 5000)), chol_mode=int(rng.integers(0, 2))))
    return out


def _engine_call(engine, sampler, c, rng_mode, arith):
    st = c["st"]
    common = dict(target_data=c["td"], n_burnin=st["n_burnin"], n_keep=st["n_keep"], rng_mode=rng_mode, seed=c["seed"], arith=arith,
                  chain_offset=c["offset"], chol_mode=c["chol_mode"], want_logp=True)
    if sampler == ol.HMC:
        return engine.hmc(c["x0"], TNAME[c["tid"]], n_leap_steps=st["n_leap_steps"], step_size=st["step_size"], precond_mat=st["precond"], **common)
    if sampler == ol.MALA:
        return engine.mala(c["x0"], TNAME[c["tid"]], step_size=st["step_size"], precond_mat=st["precond"], **common)
    return engine.rwmh(c["x0"], TNAME[c["tid"]], par_scale=st["step_size"], cov_mat=st["precond"], **common)


@pytest.mark.parametrize("sampler,name,n,seed", [(ol.HMC, "hmc", 60, 101), (ol.MALA, "mala", 40, 202), (ol.RWMH, "rwmh", 60, 303)])
def test_random_configurations(engine, oracle, sampler, name, n, seed):
    api = engine.api
    checked = 0
    for c in _configs(sampler, n, seed):
        tag = "%s d=%d C=%d target=%s %r" % (name, c["d"], c["C"], TNAME[c["tid"]], {k: v for k, v in c["st"].items() if k in ("n_burnin", "n_keep", "n_leap_steps")})
        # reference stream (MT19937 tape), kernels' reduction order
        od, oa, olp = _oracle_chains(oracle, sampler, c["tid"], c["td"], c["x0"], c["st"], c["seed"], ol.RNG_MT, ol.SUM_WARP,
                                     chain_offset=c["offset"], chol_mode=c["chol_mode"])
        r = _engine_call(engine, sampler, c, api.RNG_MT19937_TAPE, api.ARITH_STRICT)
        assert np.array_equal(r["n_accept"], oa), tag
        if c["st"]["precond"] is None:
            assert np.array_equal(r["draws"], od), tag          # bit-exact: same operations in the same order
            assert np.array_equal(r["logp"], olp), tag
        else:
            assert np.abs(r["draws"] - od).max() <= TOL, tag   # host-side factorisation order differs in the last bit
        r = _engine_call(engine, sampler, c, api.RNG_MT19937_TAPE, api.ARITH_FAST)
        assert np.abs(r["draws"] - od).max() <= TOL and np.array_equal(r["n_accept"], oa), tag
        # production stream (Philox, global chain ids)
        od, oa, _ = _oracle_chains(oracle, sampler, c["tid"], c["td"], c["x0"], c["st"], c["seed"], ol.RNG_PHILOX, ol.SUM_WARP,
                                   chain_offset=c["offset"], chol_mode=c["chol_mode"])
        r = _engine_call(engine, sampler, c, api.RNG_PHILOX, api.ARITH_FAST)
        assert np.abs(r["draws"] - od).max() <= TOL and np.array_equal(r["n_accept"], oa), tag
        checked += 1
    assert checked == n
