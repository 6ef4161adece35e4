"""Timing of the non-headline BASELINE configs on one GPU (kernel time from CUDA events inside the library)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mcmc_b200
from mcmc_b200 import api
which = sys.argv[1:] or ["c4", "c3", "sweep"]
if "c4" in which:
    rng = np.random.default_rng(11)
    d = 256
    q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    lam = np.logspace(0, 3, d)
    P = (q / lam) @ q.T; P = (P + P.T) / 2
    for C in (512, 4096):
        x0 = rng.normal(size=(C, d))
        t0 = time.time()
        r = mcmc_b200.nuts(x0, "dense_gauss", target_data=P, n_burnin=200, n_keep=200, n_adapt_draws=200, rng_mode=api.RNG_PHILOX, seed=5)
        nlf = r["n_leapfrog"].sum()
        print("C4 NUTS d=256 dense cond 1e3: %d chains x 400 draws: kernel %.1f ms, %.3e draws/s, %.3e leapfrogs/s (%.1f leapfrogs/draw), eps=%.3f, wall %.1fs"
              % (C, r["kernel_ms"], C * 400 / r["kernel_ms"] * 1e3, nlf / r["kernel_ms"] * 1e3, nlf / (C * 400), r["step_size"].mean(), time.time() - t0))
if "c3" in which:
    # BASELINE config 3: MALA, d = 1024 Bayesian linear regression on sufficient statistics, 16384 chains, 20 + 100 draws;
    # draws_out (13.4 GB) stays on the device
    import torch
    rng = np.random.default_rng(7)
    d, C, n = 1024, 16384, 4096
    X = rng.normal(size=(n, d))
    beta = np.sin(np.arange(d))
    yv = X @ beta + 0.5 * rng.normal(size=n)
    A = X.T @ X / 0.25 + np.eye(d) / 100.0
    A = (A + A.T) / 2
    b = X.T @ yv / 0.25
    eps = 0.5 / np.sqrt(np.linalg.eigvalsh(A).max())
    td = np.concatenate([A.ravel(), b])
    x0 = torch.from_numpy(rng.normal(size=(C, d)) * 0.01 + np.linalg.solve(A, b)).cuda()
    draws = torch.empty((C, 100, d), dtype=torch.float64, device="cuda")
    for _ in range(2):
        r = mcmc_b200.mala(None, "linreg", target_data=td, step_size=eps, n_burnin=20, n_keep=100, rng_mode=api.RNG_PHILOX, seed=3,
                           initial_dev_ptr=x0.data_ptr(), n_chains=C, n_dim=d, draws_dev_ptr=draws.data_ptr(),
                           stream=torch.cuda.current_stream().cuda_stream)
    print("C3 MALA d=1024 linreg: %d chains x 120 draws: kernel %.1f ms (%.3f ms/draw, %d launches), acc %.2f, gradient GEMM %.1f TFLOP/s-equivalent"
          % (C, r["kernel_ms"], r["kernel_ms"] / 120, r["kernel_launches"], r["n_accept"].mean() / 100, 2.0 * d * d * C * 121 / r["kernel_ms"] * 1e-9))
if "sweep" in which:
    for d in (32, 128, 512, 2048):
        C = 4096
        x0 = np.sin(0.37 * np.arange(C)[:, None] + 0.11 * np.arange(d)[None, :])
        for _ in range(2):   # the first launch of a kernel includes CUDA's lazy module load
            r = mcmc_b200.hmc(x0, "iso_gauss", n_leap_steps=10, step_size=0.1 * (128 / d) ** 0.25, n_burnin=100, n_keep=200, rng_mode=api.RNG_PHILOX, seed=1)
        ms = r["kernel_ms"]
        print("HMC iso d=%d: 4096 chains x 300 draws: kernel %.2f ms, %.3e draws/s, %.1f GB/s algorithmic (2*d*8 B/draw), acc %.3f"
              % (d, ms, C * 300 / ms * 1e3, C * 300 * 2 * d * 8 / ms * 1e-6, r["n_accept"].mean() / 200))
if "c5" in which:
    # BASELINE config 5: RM-HMC, Neal's funnel d = 64, 2048 chains, SoftAbs metric (alpha = 1e6, closed form), n_fp = 5, L = 5;
    # eps = 0.01: the reference's RM-HMC (Q16/Q17 semantics) accepts ~70 % there, nothing at 0.1
    rng = np.random.default_rng(5)
    d, C = 64, 2048
    x0 = rng.normal(size=(C, d)) * 0.6
    x0[:, 0] = rng.uniform(-0.5, 0.8, size=C)
    for metric, name in ((2, "funnel_softabs"), (1, "funnel_fisher")):
        for nb, nk in ((1, 2), (5, 15)):
            t0 = time.time()
            r = mcmc_b200.rmhmc(x0, "funnel", n_leap_steps=5, step_size=0.01, n_fp_steps=5, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX, seed=5,
                                metric_id=metric)
            print("C5 RM-HMC funnel d=64 (%s): %d chains x %d draws, L=5, n_fp=5: kernel %.1f ms (%.1f ms/draw), acc %.2f, finite %.3f, wall %.1fs"
                  % (name, C, nb + nk, r["kernel_ms"], r["kernel_ms"] / (nb + nk), r["n_accept"].mean() / nk, np.isfinite(r["draws"]).all(axis=(1, 2)).mean(), time.time() - t0))
