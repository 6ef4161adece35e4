"""Reduced C4 job (NUTS d=256 dense Gaussian, cond 1e3) for ncu captures: argv = n_chains n_burnin n_keep."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mcmc_b200
from mcmc_b200 import api
C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nk = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rng = np.random.default_rng(11)
d = 256
q, _ = np.linalg.qr(rng.normal(size=(d, d)))
lam = np.logspace(0, 3, d)
P = (q / lam) @ q.T; P = (P + P.T) / 2
x0 = rng.normal(size=(C, d))
r = mcmc_b200.nuts(x0, "dense_gauss", target_data=P, n_burnin=nb, n_keep=nk, n_adapt_draws=nb, rng_mode=api.RNG_PHILOX, seed=5)
nlf = r["n_leapfrog"].sum()
print("C4-small: %d chains x %d draws: kernel %.1f ms, %.3e leapfrogs/s (%.1f leapfrogs/draw)" % (C, nb + nk, r["kernel_ms"], nlf / r["kernel_ms"] * 1e3, nlf / (C * (nb + nk))))
