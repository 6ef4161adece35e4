import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mcmc_b200
os.environ["MCMCB200_DEBUG"] = "1"
d, C = 96, 300
A = np.eye(d) * 2; b = np.ones(d)
try:
    r = mcmc_b200.mala(np.zeros((C, d)), "linreg", target_data=np.concatenate([A.ravel(), b]), step_size=0.3, n_burnin=1, n_keep=2, seed=1)
    print("ok", r["kernel_launches"], r["draws"][0, -1, :3])
except Exception as e:
    print("ERR", e)
