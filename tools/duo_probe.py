"""C2 kernel time vs chains per GPU: the production one-warp-per-chain kernel against the two-warps-per-chain kernel (hmc_duo.cu)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mcmc_b200
from mcmc_b200 import api
st = torch.cuda.current_stream().cuda_stream
def run(C, d, duo, nb=100, nk=1000, L=10, reps=4):
    os.environ["MCMCB200_HMC_DUO"] = "1" if duo else "0"
    x0 = torch.from_numpy(np.sin(0.37 * np.arange(C)[:, None] + 0.11 * np.arange(d)[None, :])).cuda()
    draws = torch.empty((C, nk, d), dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(reps):
        r = mcmc_b200.hmc(None, "iso_gauss", n_leap_steps=L, step_size=0.1, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX, seed=12345,
                          initial_dev_ptr=x0.data_ptr(), n_chains=C, n_dim=d, draws_dev_ptr=draws.data_ptr(), stream=st)
        best = min(best, r["kernel_ms"])
    return best, draws.cpu().numpy()
for d, L in ((128, 10), (64, 10), (256, 10), (128, 5)):
    for C in (148, 296, 512, 1024, 1184, 2048, 4096):
        a, da = run(C, d, False, L=L)
        b, db = run(C, d, True, L=L)
        print("d=%3d L=%2d C=%5d: one warp per chain %.4f ms | two warps per chain %.4f ms (%.2fx)  identical draws: %s" % (d, L, C, a, b, a / b, np.array_equal(da, db)), flush=True)
