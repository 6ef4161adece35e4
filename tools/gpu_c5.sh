#!/bin/bash
tag=${1:-c5a}
mkdir -p gpurun_out /tmp/ncu
timeout 1200 python -m pytest tests/test_gpu_rmhmc.py -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
for i in 1 2; do timeout 300 python tools/prof_c5.py 2 6; done 2>&1 | tee gpurun_out/c5_$tag.log
MCMCB200_RMHMC_REGTILE=0 timeout 300 python tools/prof_c5.py 2 6 2>&1 | tee -a gpurun_out/c5_$tag.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rmhmc_cta_kernel -c 1 -f -o /tmp/ncu/c5 python tools/prof_c5.py 1 2 > gpurun_out/${tag}_c5.log 2>&1
python tools/ncu_summary.py /tmp/ncu/c5.ncu-rep 6144 > gpurun_out/${tag}_c5.ncu_summary.txt 2>&1
