"""n_dim <= 32: warp-per-chain kernel vs two chains per warp (hmc_half.cu), 4096 chains, the sweep's settings."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mcmc_b200
from mcmc_b200 import api
st = torch.cuda.current_stream().cuda_stream
def run(C, d, half, nb=100, nk=200, L=10, reps=4):
    os.environ["MCMCB200_HMC_HALF"] = "1" if half else "0"
    x0 = torch.from_numpy(np.sin(0.37 * np.arange(C)[:, None] + 0.11 * np.arange(d)[None, :])).cuda()
    draws = torch.empty((C, nk, d), dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(reps):
        r = mcmc_b200.hmc(None, "iso_gauss", n_leap_steps=L, step_size=0.1 * (128 / d) ** 0.25, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX, seed=12345,
                          initial_dev_ptr=x0.data_ptr(), n_chains=C, n_dim=d, draws_dev_ptr=draws.data_ptr(), stream=st)
        best = min(best, r["kernel_ms"])
    return best, draws.cpu().numpy()
for d in (32, 16, 8):
    for C in (4096, 16384):
        a, da = run(C, d, False)
        b, db = run(C, d, True)
        print("d=%2d C=%5d: warp per chain %.4f ms | two chains per warp %.4f ms (%.2fx)  identical draws: %s" % (d, C, a, b, a / b, np.array_equal(da, db)), flush=True)
