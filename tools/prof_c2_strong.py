"""C2 headline kernel on a strong-scaling shard (argv: chains per GPU, default 512 = the 8-GPU share of 4096 chains) for ncu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mcmc_b200
from mcmc_b200 import api
C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
d, nb, nk = 128, 100, 1000
x0 = torch.from_numpy(np.sin(0.37 * np.arange(C)[:, None] + 0.11 * np.arange(d)[None, :])).cuda()
draws = torch.empty((C, nk, d), dtype=torch.float64, device="cuda")
for it in range(3):
    r = mcmc_b200.hmc(None, "iso_gauss", n_leap_steps=10, step_size=0.1, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX, seed=12345,
                      initial_dev_ptr=x0.data_ptr(), n_chains=C, n_dim=d, draws_dev_ptr=draws.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    print("C=%d kernel_ms %.4f" % (C, r["kernel_ms"]))
