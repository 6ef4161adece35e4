#!/bin/bash
tag=${1:-c5a}
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests/test_gpu_rmhmc.py -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$tag.log
for i in 1 2; do timeout 300 python tools/prof_c5.py 2 6; done 2>&1 | tee gpurun_out/c5_$tag.log
python - <<'PY' 2>&1 | tee -a gpurun_out/c5_$tag.log
import numpy as np, mcmc_b200
from mcmc_b200 import api
rng = np.random.default_rng(5)
for d, C in ((96, 592), (128, 592)):
    x0 = rng.normal(size=(C, d)) * 0.6; x0[:, 0] = rng.uniform(-0.5, 0.8, size=C)
    r = mcmc_b200.rmhmc(x0, "funnel", n_leap_steps=5, step_size=0.005, n_fp_steps=5, n_burnin=1, n_keep=3, rng_mode=api.RNG_PHILOX, seed=5, metric_id=2)
    print("RM-HMC funnel SoftAbs d=%d, %d chains x 4 draws: kernel %.1f ms (%.1f ms/draw), acc %.2f, finite %.2f" % (d, C, r["kernel_ms"], r["kernel_ms"] / 4, r["n_accept"].mean() / 3, np.isfinite(r["draws"]).all(axis=(1, 2)).mean()))
PY
