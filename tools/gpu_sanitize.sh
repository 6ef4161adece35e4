#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, mcmc_b200
from mcmc_b200 import api
rng = np.random.default_rng(1)
d, C = 64, 40
q, _ = np.linalg.qr(rng.normal(size=(d, d)))
P = (q / np.logspace(0, 2, d)) @ q.T; P = (P + P.T) / 2
x0 = rng.normal(size=(C, d))
r = mcmc_b200.nuts(x0, "dense_gauss", target_data=P.ravel(), n_burnin=3, n_keep=3, n_adapt_draws=3, rng_mode=api.RNG_PHILOX, seed=5)
print("launches", r["kernel_launches"], "finite", np.isfinite(r["draws"]).all())
x0 = rng.normal(size=(24, d)) * 0.5; x0[:, 0] = 0.2
r = mcmc_b200.rmhmc(x0, "funnel", n_leap_steps=2, step_size=0.02, n_fp_steps=2, n_burnin=1, n_keep=2, rng_mode=api.RNG_PHILOX, seed=5, metric_id=2)
print("rmhmc finite", np.isfinite(r["draws"]).all())
PY
for tool in memcheck racecheck; do
  echo "== $tool persistent nuts + rmhmc"; MCMCB200_NUTS_BATCHED=1 timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | tail -12
done 2>&1 | tee gpurun_out/sanitize.log
echo "== memcheck launched rounds"; MCMCB200_NUTS_BATCHED=1 MCMCB200_NUTS_PERSIST=0 timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee -a gpurun_out/sanitize.log
