#!/bin/bash
tag=${1:-nb1}
mkdir -p gpurun_out /tmp/ncu
timeout 1200 python -m pytest tests/test_gpu_nuts_batched.py -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_$tag.log
for i in 1 2 3; do timeout 300 python tools/prof_c4.py 4096 200 200; done 2>&1 | tee gpurun_out/c4_$tag.log
MCMCB200_DEBUG=1 timeout 300 python tools/prof_c4.py 2368 100 100 2>&1 | tee -a gpurun_out/c4_$tag.log
