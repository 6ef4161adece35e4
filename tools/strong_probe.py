"""Scratch probe: C2 kernel time vs chains per GPU (strong-scaling shard sizes), d=64 half-tile kernel, bare D2H ceiling."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mcmc_b200
from mcmc_b200 import api
st = torch.cuda.current_stream().cuda_stream
def run(C, d, nb=100, nk=1000, L=10, reps=4):
    x0 = torch.from_numpy(np.sin(0.37 * np.arange(C)[:, None] + 0.11 * np.arange(d)[None, :])).cuda()
    draws = torch.empty((C, nk, d), dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(reps):
        r = mcmc_b200.hmc(None, "iso_gauss", n_leap_steps=L, step_size=0.1, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX, seed=12345,
                          initial_dev_ptr=x0.data_ptr(), n_chains=C, n_dim=d, draws_dev_ptr=draws.data_ptr(), stream=st)
        best = min(best, r["kernel_ms"])
    print("C=%5d d=%4d L=%d: kernel %.4f ms  %.3e draws/s  cycles/warp-draw@1.93GHz %.0f" % (C, d, L, best, C * (nb + nk) / best * 1e3, best * 1e-3 * 1.93e9 / (nb + nk)), flush=True)
    return best
for C in (148, 296, 512, 592, 1024, 1184, 2048, 4096):
    run(C, 128)
for C in (512, 1024, 2048, 8192):
    run(C, 64)
for L in (5, 10, 20):
    run(4096, 128, L=L)
# bare D2H ceiling
n = 4096 * 1000 * 128
dev = torch.empty(n, dtype=torch.float64, device="cuda")
host = torch.empty(n, dtype=torch.float64).pin_memory()
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); host.copy_(dev, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("D2H pinned %.2f GB in %.1f ms = %.1f GB/s" % (n * 8e-9, dt * 1e3, n * 8e-9 / dt), flush=True)
