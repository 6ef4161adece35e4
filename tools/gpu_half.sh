#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hmc.py tests/test_gpu_fuzz.py -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_half.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_half.log
timeout 600 python tools/half_probe.py 2>&1 | tee gpurun_out/half_probe.log
