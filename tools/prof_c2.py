"""Run the C2 headline kernel a few times with device-resident buffers (for ncu captures)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mcmc_b200
from mcmc_b200 import api
import oracle_lib as ol
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
C, d, nb, nk = 4096, 128, 100, 1000
x0 = torch.from_numpy(ol.c2_initial(C, d)).cuda()
draws = torch.empty((C, nk, d), dtype=torch.float64, device="cuda")
for it in range(n):
    r = mcmc_b200.hmc(None, "iso_gauss", n_leap_steps=10, step_size=0.1, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX,
                      seed=12345, arith=api.ARITH_FAST, initial_dev_ptr=x0.data_ptr(), n_chains=C, n_dim=d,
                      draws_dev_ptr=draws.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    print("kernel_ms", r["kernel_ms"])
