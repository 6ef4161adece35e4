#!/usr/bin/env python
"""Generate tests/golden/reference_golden.json from the UNMODIFIED reference (oracle/_ref/libmcmc_ref_strict.so:
/root/reference/src/{hmc,mala,nuts,rmhmc,rwmh,de}.cpp compiled against the stand-in Eigen, -O2 -ffp-contract=off).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these fixtures pin the oracle — and through it
the CUDA kernels — to outputs of the reference itself.  Floats are stored as C99 hex strings (bit-exact)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def hexlist(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def cases():
    rng = np.random.default_rng(20260925)

    def sym_pd(d, shift):
        a = rng.normal(size=(d, d))
        m = a @ a.T / d + shift * np.eye(d)
        return (m + m.T) / 2

    xs = 2 + 2 * np.sin(np.arange(100.0))
    nm = [100.0, float(xs.mean()), float(((xs - xs.mean()) ** 2).sum())]
    P6, M6 = sym_pd(6, 1.0), sym_pd(6, 0.5)
    A5, b5 = sym_pd(5, 2.0), rng.normal(size=5)
    out = [
        # SURVEY Appendix B
        dict(name="G2_hmc_d3", sampler=ol.HMC, target=ol.TGT_ISO_GAUSS, tdata=None, x0=[1, -1, 0.5], seed=1,
             st=dict(n_burnin=0, n_keep=5, n_leap_steps=10, step_size=0.1)),
        dict(name="G3_mala_d3", sampler=ol.MALA, target=ol.TGT_ISO_GAUSS, tdata=None, x0=[1, -1, 0.5], seed=1,
             st=dict(n_burnin=0, n_keep=5, step_size=0.5)),
        dict(name="G4_rmhmc_normal", sampler=ol.RMHMC, target=ol.TGT_NORMAL_MODEL, tdata=nm, x0=[3, 3], seed=1,
             st=dict(n_burnin=0, n_keep=5, n_leap_steps=1, step_size=0.2)),
        dict(name="G5_nuts_1d", sampler=ol.NUTS, target=ol.TGT_ISO_GAUSS, tdata=None, x0=[0.3], seed=3,
             st=dict(n_burnin=0, n_keep=3, step_size=0.05, n_adapt_draws=0)),
        # C2 parity-subset shape (one chain of it: chain 5, seed 12345+5)
        dict(name="C2_hmc_d128_chain5", sampler=ol.HMC, target=ol.TGT_ISO_GAUSS, tdata=None,
             x0=ol.c2_initial(1, 128, 5)[0].tolist(), seed=12350, st=dict(n_burnin=10, n_keep=20, n_leap_steps=10, step_size=0.1)),
        dict(name="hmc_dense_mass_dense_target", sampler=ol.HMC, target=ol.TGT_DENSE_GAUSS, tdata=P6.ravel().tolist(),
             x0=rng.normal(size=6).tolist(), seed=3, st=dict(n_burnin=5, n_keep=40, n_leap_steps=5, step_size=0.15, precond=M6.tolist())),
        dict(name="hmc_rejections", sampler=ol.HMC, target=ol.TGT_ISO_GAUSS, tdata=None, x0=ol.c2_initial(1, 32)[0].tolist(), seed=21,
             st=dict(n_burnin=0, n_keep=60, n_leap_steps=3, step_size=1.3)),
        dict(name="mala_d16", sampler=ol.MALA, target=ol.TGT_ISO_GAUSS, tdata=None, x0=rng.normal(size=16).tolist(), seed=11,
             st=dict(n_burnin=10, n_keep=60, step_size=0.6)),
        dict(name="mala_linreg_dense_precond", sampler=ol.MALA, target=ol.TGT_LINREG, tdata=np.concatenate([A5.ravel(), b5]).tolist(),
             x0=rng.normal(size=5).tolist(), seed=5, st=dict(n_burnin=10, n_keep=60, step_size=0.25, precond=(sym_pd(5, 0.5) / 3).tolist())),
        dict(name="nuts_d3_adapt", sampler=ol.NUTS, target=ol.TGT_DIAG_GAUSS, tdata=[1.0, 4.0, 0.25], x0=[0.1, 0.2, 0.3], seed=1,
             st=dict(n_burnin=50, n_keep=50, n_adapt_draws=50)),
        dict(name="nuts_d8_deep_trees", sampler=ol.NUTS, target=ol.TGT_ISO_GAUSS, tdata=None, x0=rng.normal(size=8).tolist(), seed=2,
             st=dict(n_burnin=0, n_keep=30, step_size=0.005, n_adapt_draws=0)),
        dict(name="rmhmc_L2", sampler=ol.RMHMC, target=ol.TGT_NORMAL_MODEL, tdata=nm, x0=[3, 3], seed=2,
             st=dict(n_burnin=10, n_keep=80, n_leap_steps=2, step_size=0.15)),
    ]
    # box constraints (algo_settings_t::vals_bound): all four bound types in one vector; +-inf = open side
    inf = float("inf")
    lo4, hi4 = [-inf, 0.0, -inf, -1.0], [inf, inf, 2.0, 1.5]
    out += [
        dict(name="hmc_box_d4", sampler=ol.HMC, target=ol.TGT_DIAG_GAUSS, tdata=[1.0, 0.5, 2.0, 1.5], x0=[0.3, 0.7, 0.4, 0.2], seed=31,
             st=dict(n_burnin=5, n_keep=40, n_leap_steps=6, step_size=0.2), lower=lo4, upper=hi4),
        dict(name="mala_box_d4", sampler=ol.MALA, target=ol.TGT_DIAG_GAUSS, tdata=[1.0, 0.5, 2.0, 1.5], x0=[0.3, 0.7, 0.4, 0.2], seed=32,
             st=dict(n_burnin=5, n_keep=40, step_size=0.35), lower=lo4, upper=hi4),
        dict(name="nuts_box_d4", sampler=ol.NUTS, target=ol.TGT_DIAG_GAUSS, tdata=[1.0, 0.5, 2.0, 1.5], x0=[0.3, 0.7, 0.4, 0.2], seed=33,
             st=dict(n_burnin=0, n_keep=30, step_size=0.15, n_adapt_draws=0), lower=lo4, upper=hi4),
        dict(name="rmhmc_box_sigma_positive", sampler=ol.RMHMC, target=ol.TGT_NORMAL_MODEL, tdata=nm, x0=[3, 3], seed=34,
             st=dict(n_burnin=5, n_keep=40, n_leap_steps=2, step_size=0.1), lower=[-inf, 0.0], upper=[inf, inf]),
    ]
    # mcmc::rwmh (src/rwmh.cpp): st.step_size carries par_scale, st.precond carries cov_mat.  Appended last so that
    # the cases above keep consuming the same numpy stream as before (their fixtures do not change).
    C6 = sym_pd(6, 0.5)
    out += [
        dict(name="rwmh_d3", sampler=ol.RWMH, target=ol.TGT_ISO_GAUSS, tdata=None, x0=[1, -1, 0.5], seed=1,
             st=dict(n_burnin=0, n_keep=8, step_size=0.5)),
        dict(name="rwmh_dense_cov_dense_target", sampler=ol.RWMH, target=ol.TGT_DENSE_GAUSS, tdata=P6.ravel().tolist(),
             x0=rng.normal(size=6).tolist(), seed=41, st=dict(n_burnin=10, n_keep=60, step_size=0.45, precond=C6.tolist())),
        dict(name="rwmh_d128", sampler=ol.RWMH, target=ol.TGT_ISO_GAUSS, tdata=None, x0=ol.c2_initial(1, 128, 7)[0].tolist(), seed=12352,
             st=dict(n_burnin=10, n_keep=40, step_size=0.2)),
        dict(name="rwmh_box_d4", sampler=ol.RWMH, target=ol.TGT_DIAG_GAUSS, tdata=[1.0, 0.5, 2.0, 1.5], x0=[0.3, 0.7, 0.4, 0.2], seed=35,
             st=dict(n_burnin=5, n_keep=50, step_size=0.6), lower=lo4, upper=hi4),
    ]
    # Neal's funnel (BASELINE config 5's density) under RM-HMC with its two registered metrics (st.metric_id: 1 = Fisher-type
    # diagonal, 2 = closed-form SoftAbs, alpha = 1e6) and under HMC; appended last like the rwmh cases
    xf = [0.3] + (0.7 * rng.normal(size=5)).tolist()
    out += [
        dict(name="rmhmc_funnel_fisher_d6", sampler=ol.RMHMC, target=ol.TGT_FUNNEL, tdata=None, x0=xf, seed=51,
             st=dict(n_burnin=3, n_keep=40, n_leap_steps=3, step_size=0.1, n_fp_steps=4, metric_id=1)),
        dict(name="rmhmc_funnel_softabs_d6", sampler=ol.RMHMC, target=ol.TGT_FUNNEL, tdata=None, x0=xf, seed=52,
             st=dict(n_burnin=3, n_keep=40, n_leap_steps=3, step_size=0.08, n_fp_steps=5, metric_id=2)),
        dict(name="hmc_funnel_d6", sampler=ol.HMC, target=ol.TGT_FUNNEL, tdata=None, x0=xf, seed=53,
             st=dict(n_burnin=3, n_keep=40, n_leap_steps=5, step_size=0.1)),
    ]
    return out


def main():
    ref = ol.Reference("strict")
    golden = dict(generator="tests/golden/make_golden.py", source="oracle/_ref/libmcmc_ref_strict.so (unmodified /root/reference sources, "
                  "stand-in Eigen, g++ -O2 -ffp-contract=off, libstdc++ <random>)", cases=[])
    golden["rng_G1"] = dict(seed=1, n_norm=4, n_unif=1, values=hexlist(ref.rng_stream(1, 4, 1)))
    for c in cases():
        st = ol.Settings(**c["st"])
        if "lower" in c:   # bounds travel as hex strings ("inf" / "-inf" for the open sides): strict JSON has no Infinity
            st["lower_bounds"], st["upper_bounds"] = c["lower"], c["upper"]
        draws, acc = ref.run_chain(c["sampler"], c["target"], c["tdata"], c["x0"], st, c["seed"])
        e = dict(c)
        if "lower" in c:
            e["lower"], e["upper"] = hexlist(c["lower"]), hexlist(c["upper"])
        e["draws_shape"] = list(draws.shape)
        e["draws_hex"] = hexlist(draws)
        e["n_accept"] = int(acc)
        golden["cases"].append(e)
    # mcmc::de (src/de.cpp): one population per case, single-threaded member loop; draws [n_keep][n_pop][d]
    inf = float("inf")
    golden["de_cases"] = []
    for c in [
        dict(name="de_iso_d3", target=ol.TGT_ISO_GAUSS, tdata=None, x0=[0.5, -0.5, 1.0], seed=11, st=dict(n_pop=12, n_burnin=5, n_keep=20)),
        dict(name="de_diag_d6_jumps", target=ol.TGT_DIAG_GAUSS, tdata=np.linspace(0.5, 2, 6).tolist(), x0=[0.1, -0.2, 0.3, 0.0, 0.5, -0.4], seed=12,
             st=dict(n_pop=20, n_burnin=15, n_keep=25, jumps=True, par_b=1e-3)),
        dict(name="de_box_d4", target=ol.TGT_DIAG_GAUSS, tdata=[1.0, 0.5, 2.0, 1.5], x0=[0.3, 0.7, 0.4, 0.2], seed=13,
             st=dict(n_pop=10, n_burnin=5, n_keep=30), lower=[-inf, 0.0, -inf, -1.0], upper=[inf, inf, 2.0, 1.5]),
    ]:
        st = ol.DeSettings(**c["st"])
        if "lower" in c:
            st["lower_bounds"], st["upper_bounds"] = c["lower"], c["upper"]
        draws, acc = ref.run_de(c["target"], c["tdata"], c["x0"], st, c["seed"])
        e = dict(c)
        if "lower" in c:
            e["lower"], e["upper"] = hexlist(c["lower"]), hexlist(c["upper"])
        e["draws_shape"] = list(draws.shape)
        e["draws_hex"] = hexlist(draws)
        e["n_accept"] = int(acc)
        golden["de_cases"].append(e)
    path = os.path.join(HERE, "reference_golden.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes,", len(golden["cases"]), "cases +", len(golden["de_cases"]), "de cases")


if __name__ == "__main__":
    main()
