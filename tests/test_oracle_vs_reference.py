"""CPU tests (no GPU): pin the oracle.

1. oracle/liboracle.so (restated samplers) reproduces, BIT FOR BIT, the committed golden vectors that
   tests/golden/make_golden.py generated from the unmodified reference (oracle/_ref) — runs everywhere.
2. Where oracle/_ref is available (this container; the GPU box gets the prebuilt .so) the oracle is compared with the
   live reference on further seeded configurations, again bit for bit.
3. The alternative oracle modes used as GPU comparators (kernel reduction order, cancelled-form MALA, Philox) are tied
   back to the reference-equivalent mode."""
import numpy as np
import pytest

import golden_util
import oracle_lib as ol


@pytest.fixture(scope="module")
def golden():
    return golden_util.load()


def test_golden_rng_stream(oracle, golden):
    """4 x rnorm then 1 x runif from mt19937_64(1) with the BaseMatrixOps semantics."""
    g = golden["rng_G1"]
    want = np.array([float.fromhex(h) for h in g["values"]])
    st = ol.Settings(n_burnin=0, n_keep=1, n_leap_steps=0, step_size=0.1)
    o = oracle.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, np.zeros(4), st, seed=1, record_tape=16)
    assert np.array_equal(o["tape"], want)
    assert np.abs(want - [-0.38683176162103994, 0.68682363917932543, -0.79514624370949216, 1.9379462044713822,
                          0.089453193644654524]).max() < 1e-16


def test_oracle_reproduces_every_golden_case(oracle, golden):
    assert len(golden["cases"]) >= 23
    for c in golden["cases"]:
        o = oracle.run_chain(c["sampler"], c["target"], c["tdata"], c["x0"], c["settings"], seed=c["seed"], rng_mode=ol.RNG_MT,
                             sum_mode=ol.SUM_SEQ, chol_mode=1, mala_exact=1)
        assert np.array_equal(o["draws"], c["draws"]), c["name"]
        assert o["n_accept"] == c["n_accept"], c["name"]


def test_appendix_b_values(golden):
    by = {c["name"]: c for c in golden["cases"]}
    g2 = by["G2_hmc_d3"]["draws"]
    assert np.abs(g2[0] - [0.21394863357965233, 0.038869637751197325, -0.40013418025888436]).max() < 1e-16
    assert np.abs(g2[4] - [-2.3607964856140051, 0.1836841693287658, 0.44563121994797955]).max() < 1e-15
    g3 = by["G3_mala_d3"]["draws"]
    assert np.abs(g3[0] - [0.68158411918948003, -0.53158818041033729, 0.039926878145253919]).max() < 1e-16
    g4 = by["G4_rmhmc_normal"]
    assert g4["n_accept"] == 2 and np.abs(g4["draws"][2] - [3.0733742689491672, 2.9473085318575425]).max() < 1e-15
    g5 = by["G5_nuts_1d"]["draws"][:, 0]
    assert np.abs(g5 - [0.29999999999999999, 0.87358199970259187, 0.86317002105720686]).max() < 1e-16


def _rand_cases():
    rng = np.random.default_rng(314)

    def sym_pd(d, shift):
        a = rng.normal(size=(d, d))
        m = a @ a.T / d + shift * np.eye(d)
        return (m + m.T) / 2

    xs = 2 + 2 * np.sin(np.arange(100.0))
    nm = [100.0, float(xs.mean()), float(((xs - xs.mean()) ** 2).sum())]
    P, M = sym_pd(7, 1.0), sym_pd(7, 0.5)
    A, b = sym_pd(7, 2.0), rng.normal(size=7)
    return [
        ("hmc iso d=70", ol.HMC, ol.TGT_ISO_GAUSS, None, rng.normal(size=70), ol.Settings(n_burnin=5, n_keep=40, n_leap_steps=7, step_size=0.3), 5),
        ("hmc diag", ol.HMC, ol.TGT_DIAG_GAUSS, np.linspace(0.5, 2, 9), rng.normal(size=9), ol.Settings(n_burnin=5, n_keep=60, n_leap_steps=5, step_size=0.4), 6),
        ("hmc dense/dense", ol.HMC, ol.TGT_DENSE_GAUSS, P.ravel(), rng.normal(size=7), ol.Settings(n_burnin=5, n_keep=60, n_leap_steps=4, step_size=0.2, precond=M), 7),
        ("hmc linreg", ol.HMC, ol.TGT_LINREG, np.concatenate([A.ravel(), b]), rng.normal(size=7), ol.Settings(n_burnin=5, n_keep=60, n_leap_steps=4, step_size=0.1), 8),
        ("hmc normal model", ol.HMC, ol.TGT_NORMAL_MODEL, nm, [3, 3], ol.Settings(n_burnin=10, n_keep=100, n_leap_steps=5, step_size=0.08), 9),
        ("hmc unstable", ol.HMC, ol.TGT_ISO_GAUSS, None, rng.normal(size=6), ol.Settings(n_burnin=0, n_keep=15, n_leap_steps=400, step_size=2.5), 10),
        ("mala iso", ol.MALA, ol.TGT_ISO_GAUSS, None, rng.normal(size=20), ol.Settings(n_burnin=5, n_keep=80, step_size=0.5), 11),
        ("mala linreg M", ol.MALA, ol.TGT_LINREG, np.concatenate([A.ravel(), b]), rng.normal(size=7), ol.Settings(n_burnin=5, n_keep=80, step_size=0.3, precond=M / 2), 12),
        ("nuts aniso adapt", ol.NUTS, ol.TGT_DIAG_GAUSS, np.exp(rng.uniform(-1, 1, 6)), rng.normal(size=6), ol.Settings(n_burnin=40, n_keep=40, n_adapt_draws=40), 13),
        ("nuts dense M", ol.NUTS, ol.TGT_DENSE_GAUSS, P.ravel(), rng.normal(size=7), ol.Settings(n_burnin=20, n_keep=30, n_adapt_draws=20, precond=M), 14),
        ("nuts depth cap 3", ol.NUTS, ol.TGT_ISO_GAUSS, None, rng.normal(size=5), ol.Settings(n_burnin=0, n_keep=40, n_adapt_draws=0, step_size=0.01, max_tree_depth=3), 15),
        ("rmhmc L3", ol.RMHMC, ol.TGT_NORMAL_MODEL, nm, [2.5, 2.5], ol.Settings(n_burnin=5, n_keep=100, n_leap_steps=3, step_size=0.1, n_fp_steps=4), 16),
        # mcmc::rwmh: step_size carries par_scale, precond carries cov_mat
        ("rwmh iso d=70", ol.RWMH, ol.TGT_ISO_GAUSS, None, rng.normal(size=70), ol.Settings(n_burnin=5, n_keep=80, step_size=0.15), 17),
        ("rwmh dense cov", ol.RWMH, ol.TGT_DENSE_GAUSS, P.ravel(), rng.normal(size=7), ol.Settings(n_burnin=5, n_keep=80, step_size=0.5, precond=M), 18),
        ("rwmh linreg", ol.RWMH, ol.TGT_LINREG, np.concatenate([A.ravel(), b]), rng.normal(size=7), ol.Settings(n_burnin=0, n_keep=80, step_size=0.3), 19),
        ("rwmh normal model", ol.RWMH, ol.TGT_NORMAL_MODEL, nm, [3, 3], ol.Settings(n_burnin=10, n_keep=100, step_size=0.1), 20),
        # Neal's funnel (BASELINE config 5) with its registered metrics: general-d RM-HMC (d^3 derivative cube, LU inverse, ...)
        ("rmhmc funnel fisher d=9", ol.RMHMC, ol.TGT_FUNNEL, None, np.concatenate([[0.2], 0.6 * rng.normal(size=8)]),
         ol.Settings(n_burnin=3, n_keep=30, n_leap_steps=3, step_size=0.08, n_fp_steps=4, metric_id=1), 21),
        ("rmhmc funnel softabs d=7", ol.RMHMC, ol.TGT_FUNNEL, None, np.concatenate([[-0.1], 0.6 * rng.normal(size=6)]),
         ol.Settings(n_burnin=3, n_keep=30, n_leap_steps=2, step_size=0.1, n_fp_steps=5, metric_id=2), 22),
        ("nuts funnel d=5", ol.NUTS, ol.TGT_FUNNEL, None, np.concatenate([[0.2], 0.6 * rng.normal(size=4)]),
         ol.Settings(n_burnin=20, n_keep=20, n_adapt_draws=20), 23),
        ("mala funnel d=12", ol.MALA, ol.TGT_FUNNEL, None, np.concatenate([[0.2], 0.6 * rng.normal(size=11)]),
         ol.Settings(n_burnin=3, n_keep=40, step_size=0.1), 24),
    ]


def test_oracle_bit_equal_to_live_reference(oracle, reference):
    for name, sampler, tid, tdata, x0, st, seed in _rand_cases():
        ref, acc = reference.run_chain(sampler, tid, tdata, x0, st, seed)
        o = oracle.run_chain(sampler, tid, tdata, x0, st, seed=seed, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_SEQ, chol_mode=1,
                             mala_exact=1)
        assert np.array_equal(o["draws"], ref), name
        assert o["n_accept"] == acc, name


def _bounded_cases():
    """Box constraints (vals_bound): every bound type, all four samplers (bounded MALA: M = I, the only form restated)."""
    rng = np.random.default_rng(2718)
    inf = np.inf
    xs = 2 + 2 * np.sin(np.arange(100.0))
    nm = [100.0, float(xs.mean()), float(((xs - xs.mean()) ** 2).sum())]
    d = 10
    lo = np.array([-inf, 0.0, -inf, -1.0, -2.0, -inf, 0.5, -inf, -3.0, -inf])
    hi = np.array([inf, inf, 2.0, 1.5, 2.0, inf, inf, 0.0, 3.0, 4.0])
    x0 = np.array([0.1, 0.4, 1.0, 0.2, -0.5, 0.3, 1.5, -0.7, 0.0, 1.0])
    w = np.exp(rng.uniform(-0.5, 0.5, d))
    B = dict(lower_bounds=lo, upper_bounds=hi)
    return [
        ("hmc box diag", ol.HMC, ol.TGT_DIAG_GAUSS, w, x0, ol.Settings(n_burnin=5, n_keep=60, n_leap_steps=5, step_size=0.1, **B), 41),
        ("hmc box iso lower only", ol.HMC, ol.TGT_ISO_GAUSS, None, np.abs(x0) + 0.1,
         ol.Settings(n_burnin=5, n_keep=60, n_leap_steps=4, step_size=0.15, lower_bounds=np.zeros(d)), 42),
        ("hmc box normal model", ol.HMC, ol.TGT_NORMAL_MODEL, nm, [3, 3],
         ol.Settings(n_burnin=10, n_keep=80, n_leap_steps=5, step_size=0.05, lower_bounds=[-inf, 0.0], upper_bounds=[inf, inf]), 43),
        ("mala box diag", ol.MALA, ol.TGT_DIAG_GAUSS, w, x0, ol.Settings(n_burnin=5, n_keep=80, step_size=0.2, **B), 44),
        ("nuts box diag", ol.NUTS, ol.TGT_DIAG_GAUSS, w, x0, ol.Settings(n_burnin=0, n_keep=30, n_adapt_draws=0, step_size=0.1, **B), 45),
        ("nuts box adapt", ol.NUTS, ol.TGT_DIAG_GAUSS, w, x0, ol.Settings(n_burnin=20, n_keep=20, n_adapt_draws=20, **B), 46),
        ("rmhmc box", ol.RMHMC, ol.TGT_NORMAL_MODEL, nm, [2.5, 2.5],
         ol.Settings(n_burnin=5, n_keep=60, n_leap_steps=2, step_size=0.1, lower_bounds=[-inf, 0.0], upper_bounds=[10.0, inf]), 47),
        ("rwmh box diag", ol.RWMH, ol.TGT_DIAG_GAUSS, w, x0, ol.Settings(n_burnin=5, n_keep=80, step_size=0.3, **B), 48),
        ("rwmh box diag cov", ol.RWMH, ol.TGT_DIAG_GAUSS, w, x0,
         ol.Settings(n_burnin=5, n_keep=80, step_size=0.3, precond=np.diag(np.linspace(0.5, 2.0, d)) + 0.1, **B), 49),
    ]


def test_oracle_bit_equal_to_live_reference_with_bounds(oracle, reference):
    for name, sampler, tid, tdata, x0, st, seed in _bounded_cases():
        ref, acc = reference.run_chain(sampler, tid, tdata, x0, st, seed)
        o = oracle.run_chain(sampler, tid, tdata, x0, st, seed=seed, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_SEQ, chol_mode=1, mala_exact=1)
        assert np.array_equal(o["draws"], ref), name
        assert o["n_accept"] == acc, name
        lo = -np.inf if st["lower_bounds"] is None else np.asarray(st["lower_bounds"])
        hi = np.inf if st["upper_bounds"] is None else np.asarray(st["upper_bounds"])
        assert np.all(ref >= lo) and np.all(ref <= hi), name


def test_comparator_modes_agree_with_reference_mode_with_bounds(oracle):
    for name, sampler, tid, tdata, x0, st, seed in _bounded_cases():
        if sampler == ol.NUTS and st["n_adapt_draws"] > 0:
            continue
        a = oracle.run_chain(sampler, tid, tdata, x0, st, seed=seed, sum_mode=ol.SUM_SEQ, mala_exact=1)
        b = oracle.run_chain(sampler, tid, tdata, x0, st, seed=seed, sum_mode=ol.SUM_WARP, mala_exact=0)
        assert np.abs(a["draws"] - b["draws"]).max() <= 1e-10, name
        assert a["n_accept"] == b["n_accept"], name


def test_comparator_modes_agree_with_reference_mode(oracle):
    """The GPU is compared with the oracle in SUM_WARP order / cancelled-form MALA; those modes must give the same
    draws as the reference-equivalent mode whenever no accept decision sits within rounding of its threshold."""
    for name, sampler, tid, tdata, x0, st, seed in _rand_cases():
        if sampler == ol.NUTS and st["n_adapt_draws"] > 0:
            continue  # see test_nuts_adaptation_amplifies_rounding
        a = oracle.run_chain(sampler, tid, tdata, x0, st, seed=seed, sum_mode=ol.SUM_SEQ, mala_exact=1)
        b = oracle.run_chain(sampler, tid, tdata, x0, st, seed=seed, sum_mode=ol.SUM_WARP, mala_exact=0)
        assert np.abs(a["draws"] - b["draws"]).max() <= 1e-10, name
        assert a["n_accept"] == b["n_accept"], name


def test_nuts_adaptation_amplifies_rounding(oracle):
    """Documents why adaptive-NUTS parity uses a looser tolerance: the reference algorithm, run twice on the CPU with
    only the summation order of its dot products changed (last-bit differences in U and K), keeps every decision but
    its draws drift apart by orders of magnitude more than without adaptation (where they stay bit-identical)."""
    rng = np.random.default_rng(5)
    w = np.exp(rng.uniform(-1.5, 1.5, size=12))
    x0 = rng.normal(size=(6, 12))
    worst_adapt, worst_fixed = 0.0, 0.0
    for c in range(6):
        st = ol.Settings(n_burnin=60, n_keep=60, n_adapt_draws=60)
        a = oracle.run_chain(ol.NUTS, ol.TGT_DIAG_GAUSS, w, x0[c], st, seed=100 + c, sum_mode=ol.SUM_SEQ)
        b = oracle.run_chain(ol.NUTS, ol.TGT_DIAG_GAUSS, w, x0[c], st, seed=100 + c, sum_mode=ol.SUM_WARP)
        assert a["n_accept"] == b["n_accept"]
        worst_adapt = max(worst_adapt, np.abs(a["draws"] - b["draws"]).max())
        st = ol.Settings(n_burnin=60, n_keep=60, n_adapt_draws=0, step_size=0.3)
        a = oracle.run_chain(ol.NUTS, ol.TGT_DIAG_GAUSS, w, x0[c], st, seed=100 + c, sum_mode=ol.SUM_SEQ)
        b = oracle.run_chain(ol.NUTS, ol.TGT_DIAG_GAUSS, w, x0[c], st, seed=100 + c, sum_mode=ol.SUM_WARP)
        worst_fixed = max(worst_fixed, np.abs(a["draws"] - b["draws"]).max())
    assert worst_fixed <= 1e-12
    assert 1e-12 < worst_adapt < 2e-5


def test_philox_normals_are_standard_normal(oracle):
    """The engine's production RNG as restated by the oracle: moments and independence of the Box-Muller output."""
    z = np.concatenate([oracle.rng_stream(ol.RNG_PHILOX, 7, c, 3, 512, 1)[:512] for c in range(300)])
    n = z.size
    assert abs(z.mean()) < 4 / np.sqrt(n)
    assert abs(z.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs((z ** 4).mean() - 3) < 4 * np.sqrt(96 / n)
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 4 / np.sqrt(n / 2)
    u = np.array([oracle.rng_stream(ol.RNG_PHILOX, 7, c, t, 2, 2)[2:] for c in range(50) for t in range(40)]).ravel()
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 4 / np.sqrt(12 * u.size)


def test_philox_stream_is_sharding_invariant(oracle):
    """Counters carry the GLOBAL chain id: a chain's draws do not depend on which shard it belongs to."""
    st = ol.Settings(n_burnin=2, n_keep=6, n_leap_steps=3, step_size=0.2)
    x0 = ol.c2_initial(1, 10, 77)[0]
    a = oracle.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, seed=5, rng_mode=ol.RNG_PHILOX, chain_id=77)
    b = oracle.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, seed=5, rng_mode=ol.RNG_PHILOX, chain_id=78)
    assert not np.array_equal(a["draws"], b["draws"])
    c = oracle.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, seed=5, rng_mode=ol.RNG_PHILOX, chain_id=77)
    assert np.array_equal(a["draws"], c["draws"])


def test_funnel_softabs_metric_closed_form(oracle):
    """The closed-form SoftAbs metric of Neal's funnel (alpha = 1e6: G = |H| up to 1e-6) against a numerical
    eigendecomposition of the Hessian, and its derivative cube against central differences."""
    rng = np.random.default_rng(1)
    for d in (2, 3, 6, 9, 17):
        x = rng.normal(size=d) * 0.8
        x[0] = rng.uniform(-1.0, 1.0)
        G, dG = oracle.metric(ol.TGT_FUNNEL, 2, None, x)
        v, xs = x[0], x[1:]
        e, S = np.exp(-v), (xs ** 2).sum()
        H = np.zeros((d, d))
        H[0, 0] = -1 / 9 - 0.5 * e * S
        H[0, 1:] = H[1:, 0] = e * xs
        H[1:, 1:] = -e * np.eye(d - 1)
        lam, Q = np.linalg.eigh(H)
        assert np.abs(G - (Q * np.abs(lam)) @ Q.T).max() <= 1e-13
        assert np.abs(G - G.T).max() <= 1e-15 and np.linalg.eigvalsh(G).min() > 0
        for k in range(d):
            h = 1e-6
            xp, xm = x.copy(), x.copy()
            xp[k] += h
            xm[k] -= h
            Gp, _ = oracle.metric(ol.TGT_FUNNEL, 2, None, xp, False)
            Gm, _ = oracle.metric(ol.TGT_FUNNEL, 2, None, xm, False)
            assert np.abs((Gp - Gm) / (2 * h) - dG[k]).max() <= 2e-8 * max(1.0, np.abs(dG[k]).max())
        # the Fisher-type metric: diagonal, derivative only with respect to v
        G1, dG1 = oracle.metric(ol.TGT_FUNNEL, 1, None, x)
        assert np.allclose(np.diag(G1), [1 / 9 + (d - 1) / 2] + [e] * (d - 1)) and np.abs(G1 - np.diag(np.diag(G1))).max() == 0
        assert np.abs(dG1[1:]).max() == 0 and np.allclose(np.diag(dG1[0]), [0] + [-e] * (d - 1))


# ---- mcmc::de: the oracle restatement and the tape a device kernel will consume ---------------------

def _de_cases():
    rng = np.random.default_rng(9)
    inf = np.inf
    return [
        ("iso d=3", ol.TGT_ISO_GAUSS, None, [0.5, -0.5, 1.0], ol.DeSettings(n_pop=12, n_burnin=5, n_keep=20), 11),
        ("diag d=6 jumps", ol.TGT_DIAG_GAUSS, np.linspace(0.5, 2, 6), rng.normal(size=6), ol.DeSettings(n_pop=20, n_burnin=15, n_keep=25, jumps=True, par_b=1e-3), 12),
        ("funnel d=5 initial bounds", ol.TGT_FUNNEL, None, np.zeros(5),
         ol.DeSettings(n_pop=16, n_burnin=10, n_keep=20, initial_lb=-np.ones(5), initial_ub=np.ones(5)), 13),
        ("box diag d=4", ol.TGT_DIAG_GAUSS, [1.0, 0.5, 2.0, 1.5], [0.3, 0.7, 0.4, 0.2],
         ol.DeSettings(n_pop=10, n_burnin=5, n_keep=30, lower_bounds=[-inf, 0.0, -inf, -1.0], upper_bounds=[inf, inf, 2.0, 1.5]), 14),
        ("normal model", ol.TGT_NORMAL_MODEL, [100.0, 2.0, 400.0], [2.5, 2.5], ol.DeSettings(n_pop=8, n_burnin=10, n_keep=40, par_b=1e-2), 15),
        ("dense d=7 default population", ol.TGT_DENSE_GAUSS, (lambda a: (a @ a.T / 7 + np.eye(7)).ravel())(rng.normal(size=(7, 7))), rng.normal(size=7),
         ol.DeSettings(n_burnin=3, n_keep=6), 16),
    ]


def test_de_oracle_bit_equal_to_live_reference_and_tape_replay(oracle, reference):
    """The restated differential-evolution sampler reproduces the unmodified src/de.cpp (single-threaded member loop) bit for bit;
    the recorded variate tape (what a device kernel will consume) replays to the same draws."""
    for name, tid, tdata, x0, st, seed in _de_cases():
        ref, acc = reference.run_de(tid, tdata, x0, st, seed)
        o = oracle.run_de(tid, tdata, x0, st, seed=seed, record_tape=2_000_000)
        assert np.array_equal(o["draws"], ref), name
        assert o["n_accept"] == acc, name
        d, n_pop, n_total = len(x0), st["n_pop"], st["n_burnin"] + st["n_keep"]
        assert o["tape_used"] == n_pop * d + n_total * n_pop * (d + 3), name      # static layout: a tape-driven kernel is possible
        rp = oracle.run_de(tid, tdata, x0, st, rng_mode=ol.RNG_TAPE, tape=o["tape"])
        assert np.array_equal(rp["draws"], o["draws"]) and rp["n_accept"] == o["n_accept"], name
        tp = o["tape"][n_pop * d:].reshape(n_total, n_pop, d + 3)
        i = np.arange(n_pop)[None, :]
        assert ((tp[..., 0] != i) & (tp[..., 1] != i) & (tp[..., 0] != tp[..., 1])).all(), name   # c1, c2, i pairwise distinct
        assert (np.abs(tp[..., 2:2 + d]) < st["par_b"]).all() and ((tp[..., -1] > 0) & (tp[..., -1] < 1)).all(), name


def test_de_golden_cases(oracle, golden):
    assert len(golden["de_cases"]) == 3
    for c in golden["de_cases"]:
        st = ol.DeSettings(**c["st"])
        if "lower" in c:
            st["lower_bounds"] = np.array([float.fromhex(h) for h in c["lower"]])
            st["upper_bounds"] = np.array([float.fromhex(h) for h in c["upper"]])
        want = np.array([float.fromhex(h) for h in c["draws_hex"]]).reshape(c["draws_shape"])
        o = oracle.run_de(c["target"], c["tdata"], c["x0"], st, seed=c["seed"])
        assert np.array_equal(o["draws"], want), c["name"]
        assert o["n_accept"] == c["n_accept"], c["name"]


def _swept_cases(n_cases=72):
    """A seeded sweep over the settings space (dimension, trajectory length, step size, draw counts, target family,
    identity / dense mass, open / half-open / closed boxes): the oracle is pinned to the unmodified reference well away
    from the hand-picked cases above."""
    rng = np.random.default_rng(20261017)
    cases = []
    for k in range(n_cases):
        sampler = [ol.HMC, ol.MALA, ol.NUTS, ol.RWMH][k % 4]
        d = int(rng.integers(1, 41))
        fam = int(rng.integers(0, 4))
        if fam == 0:
            tid, tdata = ol.TGT_ISO_GAUSS, None
        elif fam == 1:
            tid, tdata = ol.TGT_DIAG_GAUSS, np.exp(rng.uniform(-1, 1, d))
        else:
            a = rng.normal(size=(d, d))
            P = a @ a.T / d + (0.5 + rng.uniform()) * np.eye(d)
            P = (P + P.T) / 2
            if fam == 2:
                tid, tdata = ol.TGT_DENSE_GAUSS, P.ravel()
            else:
                tid, tdata = ol.TGT_LINREG, np.concatenate([P.ravel(), rng.normal(size=d)])
        kw = dict(n_burnin=int(rng.integers(0, 8)), n_keep=int(rng.integers(1, 40)))
        eps = float(np.exp(rng.uniform(np.log(0.02), np.log(0.6))) / d ** 0.25)
        if sampler == ol.HMC:
            kw.update(n_leap_steps=int(rng.integers(1, 13)), step_size=eps)
        elif sampler == ol.NUTS:
            adapt = int(rng.integers(0, 2)) * kw["n_burnin"]
            kw.update(step_size=eps, n_adapt_draws=adapt, max_tree_depth=int(rng.integers(1, 8)))
        else:
            kw.update(step_size=eps)
        x0 = rng.normal(size=d)
        box = int(rng.integers(0, 3))
        dense_m = rng.uniform() < 0.35 and not (sampler == ol.MALA and box)   # bounded MALA: M = I is the form restated
        if dense_m:
            m = rng.normal(size=(d, d))
            M = m @ m.T / d + np.eye(d)
            kw["precond"] = (M + M.T) / 2
        if box:
            lo = np.where(rng.uniform(size=d) < 0.5, x0 - rng.uniform(0.2, 3.0, d), -np.inf)
            hi = np.where(rng.uniform(size=d) < 0.5, x0 + rng.uniform(0.2, 3.0, d), np.inf)
            if box == 2:
                hi = np.full(d, np.inf)
            kw.update(lower_bounds=lo, upper_bounds=hi)
        cases.append(("sweep %d: sampler %d family %d d=%d box=%d denseM=%d" % (k, sampler, fam, d, box, int(dense_m)), sampler, tid, tdata, x0,
                      ol.Settings(**kw), 1000 + k))
    return cases


def test_oracle_bit_equal_to_live_reference_on_a_seeded_sweep(oracle, reference):
    """Bit-equal draw for draw over 240 seeded settings, including the chains the sweep drives unstable.  Those pin one more
    piece of reference semantics (oracle.cpp Ctx::dense_jac): the reference multiplies by its diagonal operators — the
    inverse-Jacobian matrix and the identity mass matrix — as FULL d x d matrices (src/hmc.cpp:57-59,122,160,171,184), so one
    non-finite momentum / gradient element turns every other row into NaN (0 * inf); in a bounded chain inv_transform maps
    those back to finite coordinates and the proposal can be accepted (accept test on a NaN energy: std::min(0.01, NaN) =
    0.01, src/hmc.cpp:188).  Without dense_jacobian=1 the restatement keeps the element-wise products and differs exactly
    there (the second half of this test shows it is ONLY there)."""
    n_moved = n_finite = n_elementwise_differs = 0
    for name, sampler, tid, tdata, x0, st, seed in _swept_cases(240):
        ref, acc = reference.run_chain(sampler, tid, tdata, x0, st, seed)
        kw = dict(seed=seed, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_SEQ, chol_mode=1, mala_exact=1)
        o = oracle.run_chain(sampler, tid, tdata, x0, st, dense_jacobian=1, **kw)
        assert np.array_equal(o["draws"], ref, equal_nan=True), name
        assert o["n_accept"] == acc, name
        fin = bool(np.isfinite(ref).all())
        n_finite += int(fin)
        n_moved += int(fin and acc > 0)
        e = oracle.run_chain(sampler, tid, tdata, x0, st, **kw)   # element-wise products (what the device kernels do)
        same = np.array_equal(e["draws"], ref, equal_nan=True) and e["n_accept"] == acc
        if not same:
            assert st["lower_bounds"] is not None, name           # only bounded chains ...
            t = int(np.argmax([not np.array_equal(e["draws"][k], ref[k], equal_nan=True) for k in range(len(ref))]))
            at_bound = np.isclose(ref[t], st["lower_bounds"], rtol=0, atol=1e-12) | np.isclose(ref[t], st["upper_bounds"], rtol=0, atol=1e-12)
            assert at_bound.any() or not np.isfinite(ref[t]).all(), name   # ... whose trajectory left the finite range
            n_elementwise_differs += 1
    assert n_finite >= 200 and n_moved >= 170, (n_finite, n_moved)   # the sweep is not a collection of stuck or diverged chains
    assert 1 <= n_elementwise_differs <= 12, n_elementwise_differs


def test_rmhmc_oracle_bit_equal_to_live_reference_on_a_seeded_sweep(oracle, reference):
    """mcmc::rmhmc over 120 seeded settings (Normal model with its Fisher metric; Neal's funnel d = 2..12 with both registered
    metrics; trajectory lengths 1-5, 1-6 fixed-point steps, steps 0.01-0.4, a third of the cases with box constraints)."""
    rng = np.random.default_rng(777)
    xs = 2 + 2 * np.sin(np.arange(100.0))
    nm = [100.0, float(xs.mean()), float(((xs - xs.mean()) ** 2).sum())]
    n_moved = n_finite = 0
    for k in range(120):
        if k % 3 == 0:
            tid, tdata, x0, mid = ol.TGT_NORMAL_MODEL, nm, rng.uniform(1.5, 4, 2), 0
        else:
            d = int(rng.integers(2, 13))
            tid, tdata, mid = ol.TGT_FUNNEL, None, 1 + k % 2
            x0 = np.concatenate([[rng.uniform(-0.5, 0.5)], 0.6 * rng.normal(size=d - 1)])
        kw = dict(n_burnin=int(rng.integers(0, 4)), n_keep=int(rng.integers(1, 25)), n_leap_steps=int(rng.integers(1, 6)),
                  step_size=float(np.exp(rng.uniform(np.log(0.01), np.log(0.4)))), n_fp_steps=int(rng.integers(1, 7)), metric_id=mid)
        if rng.uniform() < 0.3:
            dd = len(x0)
            lo = np.where(rng.uniform(size=dd) < 0.5, x0 - rng.uniform(0.2, 3, dd), -np.inf)
            hi = np.where(rng.uniform(size=dd) < 0.5, x0 + rng.uniform(0.2, 3, dd), np.inf)
            if tid == ol.TGT_NORMAL_MODEL:
                lo[1] = max(lo[1], 0.0) if np.isfinite(lo[1]) else 0.0   # sigma > 0
            kw.update(lower_bounds=lo, upper_bounds=hi)
        st = ol.Settings(**kw)
        ref, acc = reference.run_chain(ol.RMHMC, tid, tdata, x0, st, 5000 + k)
        o = oracle.run_chain(ol.RMHMC, tid, tdata, x0, st, seed=5000 + k, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_SEQ, chol_mode=1, dense_jacobian=1)
        assert np.array_equal(o["draws"], ref, equal_nan=True), (k, kw)
        assert o["n_accept"] == acc, (k, kw)
        n_finite += int(np.isfinite(ref).all())
        n_moved += int(acc > 0)
    assert n_finite >= 100 and n_moved >= 90, (n_finite, n_moved)


def test_de_oracle_bit_equal_to_live_reference_on_a_seeded_sweep(oracle, reference):
    """mcmc::de over 150 seeded settings (n_dim 1-16, four target families, populations of 3-29 members, jumps on / off,
    par_b 1e-5..1e-1, par_gamma_jump 0.5-2.5, default / explicit initial box / box constraints): src/de.cpp:30-271."""
    rng = np.random.default_rng(4242)
    n_moved = 0
    for k in range(150):
        d, fam = int(rng.integers(1, 17)), int(rng.integers(0, 4))
        if fam == 0:
            tid, tdata = ol.TGT_ISO_GAUSS, None
        elif fam == 1:
            tid, tdata = ol.TGT_DIAG_GAUSS, np.exp(rng.uniform(-1, 1, d))
        elif fam == 2:
            a = rng.normal(size=(d, d))
            P = a @ a.T / d + np.eye(d)
            tid, tdata = ol.TGT_DENSE_GAUSS, ((P + P.T) / 2).ravel()
        else:
            d = max(d, 2)
            tid, tdata = ol.TGT_FUNNEL, None
        x0 = rng.normal(size=d)
        kw = dict(n_pop=int(rng.integers(3, 30)), n_burnin=int(rng.integers(0, 10)), n_keep=int(rng.integers(1, 20)), jumps=bool(rng.integers(0, 2)),
                  par_b=float(10 ** rng.uniform(-5, -1)), par_gamma_jump=float(rng.uniform(0.5, 2.5)))
        m = int(rng.integers(0, 3))
        if m == 1:
            kw.update(initial_lb=x0 - rng.uniform(0.1, 2, d), initial_ub=x0 + rng.uniform(0.1, 2, d))
        if m == 2:
            kw.update(lower_bounds=np.where(rng.uniform(size=d) < 0.5, x0 - rng.uniform(0.6, 3, d), -np.inf),
                      upper_bounds=np.where(rng.uniform(size=d) < 0.5, x0 + rng.uniform(0.6, 3, d), np.inf))
        st = ol.DeSettings(**kw)
        ref, acc = reference.run_de(tid, tdata, x0, st, 9000 + k)
        o = oracle.run_de(tid, tdata, x0, st, seed=9000 + k)
        assert np.array_equal(o["draws"], ref, equal_nan=True), (k, kw)
        assert o["n_accept"] == acc, (k, kw)
        n_moved += int(acc > 0)
    assert n_moved >= 140


def _harsh_cases(n_cases):
    """Like _swept_cases but aimed at the edges: step sizes up to 3 (most trajectories overflow), trajectories up to 29 steps,
    initial points up to 5 sigma out, boxes as tight as 0.05, and a dense mass matrix also for bounded MALA (the combination
    the device path refuses, §4.6)."""
    rng = np.random.default_rng(99)
    out = []
    for k in range(n_cases):
        sampler = [ol.HMC, ol.MALA, ol.NUTS, ol.RWMH][k % 4]
        d, fam = int(rng.integers(1, 25)), int(rng.integers(0, 4))
        if fam == 0:
            tid, tdata = ol.TGT_ISO_GAUSS, None
        elif fam == 1:
            tid, tdata = ol.TGT_DIAG_GAUSS, np.exp(rng.uniform(-2, 2, d))
        else:
            a = rng.normal(size=(d, d))
            P = a @ a.T / d + (0.2 + rng.uniform()) * np.eye(d)
            P = (P + P.T) / 2
            tid, tdata = (ol.TGT_DENSE_GAUSS, P.ravel()) if fam == 2 else (ol.TGT_LINREG, np.concatenate([P.ravel(), rng.normal(size=d)]))
        kw = dict(n_burnin=int(rng.integers(0, 8)), n_keep=int(rng.integers(1, 30)))
        eps = float(np.exp(rng.uniform(np.log(0.02), np.log(3.0))))
        if sampler == ol.HMC:
            kw.update(n_leap_steps=int(rng.integers(1, 30)), step_size=eps)
        elif sampler == ol.NUTS:
            kw.update(step_size=eps, n_adapt_draws=int(rng.integers(0, 2)) * kw["n_burnin"], max_tree_depth=int(rng.integers(1, 9)))
        else:
            kw.update(step_size=eps)
        x0 = rng.normal(size=d) * rng.choice([0.3, 1, 5])
        box = int(rng.integers(0, 3))
        if rng.uniform() < 0.35:
            m = rng.normal(size=(d, d))
            M = m @ m.T / d + np.eye(d)
            kw["precond"] = (M + M.T) / 2
        if box:
            lo = np.where(rng.uniform(size=d) < 0.5, x0 - rng.uniform(0.05, 3, d), -np.inf)
            hi = np.where(rng.uniform(size=d) < 0.5, x0 + rng.uniform(0.05, 3, d), np.inf)
            if box == 2:
                hi = np.full(d, np.inf)
            kw.update(lower_bounds=lo, upper_bounds=hi)
        out.append((k, sampler, tid, tdata, x0, ol.Settings(**kw), 20000 + k))
    return out


def test_oracle_bit_equal_to_live_reference_at_the_edges(oracle, reference):
    """1200 seeded settings of which about a tenth overflow: the literal mode of the restatement (dense_jacobian=1: the
    reference's diagonal operators as full matrices — inv_jacobian_adjust, eye as mass matrix, chol(J) and (J M) eps^2 in
    bounded MALA, src/mala.cpp:111-119,155-157, mala.ipp:55-56) reproduces the unmodified reference bit for bit, NaN for NaN."""
    n_nonfinite = 0
    for k, sampler, tid, tdata, x0, st, seed in _harsh_cases(1200):
        ref, acc = reference.run_chain(sampler, tid, tdata, x0, st, seed)
        o = oracle.run_chain(sampler, tid, tdata, x0, st, seed=seed, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_SEQ, chol_mode=1, mala_exact=1, dense_jacobian=1)
        assert np.array_equal(o["draws"], ref, equal_nan=True), (k, sampler, dict(st))
        assert o["n_accept"] == acc, (k, sampler, dict(st))
        n_nonfinite += int(not np.isfinite(ref).all())
    assert 60 <= n_nonfinite <= 400, n_nonfinite


def test_comparator_modes_track_the_reference_over_the_seeded_sweep(oracle, reference):
    """Second link of the parity chain over the sweep: the oracle in the mode the CUDA kernels are tested against (warp-butterfly
    reduction order, cancelled MALA proposal ratio, element-wise Jacobian products) stays within 1e-10 of the unmodified
    reference — same accept counts — on every sweep case that stays finite (observed: <= 3e-13)."""
    worst, n = 0.0, 0
    for name, sampler, tid, tdata, x0, st, seed in _swept_cases(240):
        ref, acc = reference.run_chain(sampler, tid, tdata, x0, st, seed)
        if not np.isfinite(ref).all():
            continue
        o = oracle.run_chain(sampler, tid, tdata, x0, st, seed=seed, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_WARP, chol_mode=1, mala_exact=0)
        assert o["n_accept"] == acc, name
        err = float(np.abs(o["draws"] - ref).max() / max(1.0, np.abs(ref).max()))
        assert err <= 1e-10, (name, err)
        worst, n = max(worst, err), n + 1
    assert n >= 225


def test_rmhmc_oracle_bit_equal_to_live_reference_at_the_edges(oracle, reference):
    """mcmc::rmhmc, 300 seeded settings aimed at the edges (steps up to 2, 0-7 fixed-point iterations, up to 8 leapfrog steps,
    funnel d up to 16 started up to 3 sigma out, tight boxes): about a quarter of the chains go non-finite — the reference's
    fixed-point iterations diverge and it accepts NaN energies (src/rmhmc.cpp:199-272) — and the restatement follows bit for bit."""
    rng = np.random.default_rng(31337)
    xs = 2 + 2 * np.sin(np.arange(100.0))
    nm = [100.0, float(xs.mean()), float(((xs - xs.mean()) ** 2).sum())]
    n_nonfinite = 0
    for k in range(300):
        if k % 3 == 0:
            tid, tdata, x0, mid = ol.TGT_NORMAL_MODEL, nm, rng.uniform(0.5, 6, 2), 0
        else:
            d = int(rng.integers(2, 17))
            tid, tdata, mid = ol.TGT_FUNNEL, None, 1 + k % 2
            x0 = np.concatenate([[rng.uniform(-2, 2)], rng.choice([0.3, 1, 3]) * rng.normal(size=d - 1)])
        kw = dict(n_burnin=int(rng.integers(0, 4)), n_keep=int(rng.integers(1, 25)), n_leap_steps=int(rng.integers(1, 9)),
                  step_size=float(np.exp(rng.uniform(np.log(0.01), np.log(2.0)))), n_fp_steps=int(rng.integers(0, 8)), metric_id=mid)
        if rng.uniform() < 0.4:
            dd = len(x0)
            lo = np.where(rng.uniform(size=dd) < 0.5, x0 - rng.uniform(0.05, 3, dd), -np.inf)
            hi = np.where(rng.uniform(size=dd) < 0.5, x0 + rng.uniform(0.05, 3, dd), np.inf)
            if tid == ol.TGT_NORMAL_MODEL:
                lo[1] = max(lo[1], 0.0) if np.isfinite(lo[1]) else 0.0
            kw.update(lower_bounds=lo, upper_bounds=hi)
        st = ol.Settings(**kw)
        ref, acc = reference.run_chain(ol.RMHMC, tid, tdata, x0, st, 70000 + k)
        o = oracle.run_chain(ol.RMHMC, tid, tdata, x0, st, seed=70000 + k, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_SEQ, chol_mode=1, dense_jacobian=1)
        assert np.array_equal(o["draws"], ref, equal_nan=True), (k, kw)
        assert o["n_accept"] == acc, (k, kw)
        n_nonfinite += int(not np.isfinite(ref).all())
    assert 30 <= n_nonfinite <= 150, n_nonfinite


def test_nuts_oracle_bit_equal_to_live_reference_on_long_adaptive_runs(oracle, reference):
    """mcmc::nuts, 100 seeded settings with up to 150 adaptive + 60 kept draws, trees up to depth 10, n_dim up to 60, five
    target families (condition numbers up to ~1e2 for the diagonal family), random dual-averaging constants, dense mass and box
    constraints in a third of the cases each — about half a million leapfrog steps (include/mcmc/nuts.ipp:30-241,
    src/nuts.cpp:160-332), draw for draw bit-equal."""
    rng = np.random.default_rng(5150)
    n_leapfrog = 0
    for k in range(100):
        d, fam = int(rng.integers(1, 61)), int(rng.integers(0, 5))
        if fam == 0:
            tid, tdata = ol.TGT_ISO_GAUSS, None
        elif fam == 1:
            tid, tdata = ol.TGT_DIAG_GAUSS, np.exp(rng.uniform(-2.5, 2.5, d))
        elif fam in (2, 3):
            a = rng.normal(size=(d, d))
            P = a @ a.T / d + (0.05 + rng.uniform()) * np.eye(d)
            P = (P + P.T) / 2
            tid, tdata = (ol.TGT_DENSE_GAUSS, P.ravel()) if fam == 2 else (ol.TGT_LINREG, np.concatenate([P.ravel(), rng.normal(size=d)]))
        else:
            d = max(2, min(d, 12))
            tid, tdata = ol.TGT_FUNNEL, None
        nb = int(rng.integers(0, 150))
        kw = dict(n_burnin=nb, n_keep=int(rng.integers(1, 60)), n_adapt_draws=int(rng.choice([0, nb, nb // 2, nb + 20])),
                  step_size=float(np.exp(rng.uniform(np.log(0.005), np.log(1.0)))), max_tree_depth=int(rng.integers(1, 11)),
                  target_accept_rate=float(rng.uniform(0.4, 0.9)), gamma_val=float(rng.uniform(0.02, 0.2)), t0_val=float(rng.uniform(5, 20)),
                  kappa_val=float(rng.uniform(0.6, 0.9)))
        x0 = rng.normal(size=d)
        if rng.uniform() < 0.3:
            m = rng.normal(size=(d, d))
            M = m @ m.T / d + np.eye(d)
            kw["precond"] = (M + M.T) / 2
        if rng.uniform() < 0.3:
            kw.update(lower_bounds=np.where(rng.uniform(size=d) < 0.5, x0 - rng.uniform(0.2, 3, d), -np.inf),
                      upper_bounds=np.where(rng.uniform(size=d) < 0.5, x0 + rng.uniform(0.2, 3, d), np.inf))
        st = ol.Settings(**kw)
        ref, acc = reference.run_chain(ol.NUTS, tid, tdata, x0, st, 81000 + k)
        o = oracle.run_chain(ol.NUTS, tid, tdata, x0, st, seed=81000 + k, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_SEQ, chol_mode=1, dense_jacobian=1)
        assert np.array_equal(o["draws"], ref, equal_nan=True), (k, kw)
        assert o["n_accept"] == acc, (k, kw)
        n_leapfrog += o["n_leapfrog"]
    assert n_leapfrog >= 200_000, n_leapfrog


def test_the_only_comparator_deviation_is_the_bounded_overflow_regime(oracle, reference):
    """Over the edge sweep's chains whose OUTPUT is finite (bounded MALA with a dense mass excluded: refused on the device), the
    comparator mode — warp-order reductions, cancelled MALA ratio, element-wise Jacobian / identity-mass products: what the CUDA
    kernels implement and are tested against — tracks the unmodified reference to 1e-10 with identical accept counts, EXCEPT for a
    handful of bounded HMC chains that overflowed on the way (inv_transform brings the stored draws back to finite values, and
    the reference accepts those NaN-energy proposals where the element-wise form rejects them).  Switching only the literal
    full-matrix products on (dense_jacobian=1, still warp order) removes every one of those differences: the deviation
    stated in DESIGN §4.6 is the whole deviation."""
    n = n_dev = 0
    for k, sampler, tid, tdata, x0, st, seed in _harsh_cases(1200):
        ref, acc = reference.run_chain(sampler, tid, tdata, x0, st, seed)
        if not np.isfinite(ref).all():
            continue
        if sampler == ol.MALA and st["precond"] is not None and st["lower_bounds"] is not None:
            continue
        kw = dict(seed=seed, rng_mode=ol.RNG_MT, sum_mode=ol.SUM_WARP, chol_mode=1, mala_exact=0)
        o = oracle.run_chain(sampler, tid, tdata, x0, st, **kw)
        n += 1
        scale = max(1.0, float(np.abs(ref).max()))
        if o["n_accept"] == acc and float(np.abs(o["draws"] - ref).max()) / scale <= 1e-10:
            continue
        n_dev += 1
        assert st["lower_bounds"] is not None, k                       # only chains with box constraints ...
        lit = oracle.run_chain(sampler, tid, tdata, x0, st, dense_jacobian=1, **kw)
        assert lit["n_accept"] == acc and float(np.abs(lit["draws"] - ref).max()) / scale <= 1e-10, k   # ... and only through the full-matrix products
    assert n >= 1000 and 1 <= n_dev <= 12, (n, n_dev)
