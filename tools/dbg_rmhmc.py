"""Scratch: CTA vs cube RM-HMC kernels on the funnel, small n_dim — where do they part?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mcmc_b200
from mcmc_b200 import api
np.set_printoptions(precision=17, linewidth=200)
for d, L, eps, metric in ((2, 2, 0.1, 2), (3, 3, 0.1, 2), (2, 2, 0.1, 1)):
    rng = np.random.default_rng(50 + metric)
    x0 = rng.normal(size=(12, d)) * 0.6
    x0[:, 0] = rng.uniform(-0.5, 0.8, size=12)
    kw = dict(n_leap_steps=L, step_size=eps, n_fp_steps=4, n_burnin=0, n_keep=12, want_logp=True, metric_id=metric, rng_mode=api.RNG_PHILOX, seed=77)
    a = mcmc_b200.rmhmc(x0, "funnel", **kw)
    os.environ["MCMCB200_RMHMC_CTA"] = "0"
    b = mcmc_b200.rmhmc(x0, "funnel", **kw)
    del os.environ["MCMCB200_RMHMC_CTA"]
    print("== d=%d metric=%d" % (d, metric))
    for c in range(12):
        diff = np.abs(a["draws"][c] - b["draws"][c]).max(axis=1)
        bad = np.where(~(diff <= 1e-10))[0]
        if len(bad):
            t = bad[0]
            print("chain %d first bad draw %d diff %.3e acc cta/cube %d/%d" % (c, t, diff[t], a["n_accept"][c], b["n_accept"][c]))
            lo = max(0, t - 2)
            print("  cta  draws", a["draws"][c][lo:t + 1].tolist(), "logp", a["logp"][c][lo:t + 1].tolist())
            print("  cube draws", b["draws"][c][lo:t + 1].tolist(), "logp", b["logp"][c][lo:t + 1].tolist())
