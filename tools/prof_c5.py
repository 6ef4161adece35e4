"""C5 job (RM-HMC, Neal's funnel d=64, SoftAbs metric, 2048 chains, L=5, n_fp=5) for ncu captures: argv = n_burnin n_keep."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mcmc_b200
from mcmc_b200 import api
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(5)
d, C = 64, 2048
x0 = rng.normal(size=(C, d)) * 0.6
x0[:, 0] = rng.uniform(-0.5, 0.8, size=C)
r = mcmc_b200.rmhmc(x0, "funnel", n_leap_steps=5, step_size=0.01, n_fp_steps=5, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX, seed=5, metric_id=2)
print("C5: %d chains x %d draws: kernel %.1f ms (%.1f ms/draw), acc %.2f" % (C, nb + nk, r["kernel_ms"], r["kernel_ms"] / (nb + nk), r["n_accept"].mean() / nk))
