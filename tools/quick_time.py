"""Scratch timing of the C2 headline kernel (device-resident buffers via torch), printing kernel ms and draws/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import mcmc_b200
from mcmc_b200 import api
import oracle_lib as ol
C, d, nb, nk = 4096, 128, 100, 1000
x0 = torch.from_numpy(ol.c2_initial(C, d)).cuda()
draws = torch.empty((C, nk, d), dtype=torch.float64, device="cuda")
for arith in (api.ARITH_FAST, api.ARITH_STRICT):
    for it in range(4):
        torch.cuda.synchronize()
        r = mcmc_b200.hmc(None, "iso_gauss", n_leap_steps=10, step_size=0.1, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX,
                          seed=12345, arith=arith, initial_dev_ptr=x0.data_ptr(), n_chains=C, n_dim=d, draws_dev_ptr=draws.data_ptr(),
                          stream=torch.cuda.current_stream().cuda_stream)
        ms = r["kernel_ms"]
        print("arith=%d iter %d: kernel %.3f ms -> %.3e draws/s, %.3e leapfrog/s, acc=%.4f" % (arith, it, ms, C*(nb+nk)/ms*1e3, C*(nb+nk)*10/ms*1e3, r["n_accept"].mean()/nk))
print("mean", draws[:, 500:].mean().item(), "var", draws[:, 500:].var().item())
