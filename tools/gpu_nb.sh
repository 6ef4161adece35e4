#!/bin/bash
tag=${1:-nb1}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nuts_batched.py -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$tag.log
for C in 512 1024 1184; do
  for NW in 8 16; do echo "C=$C NW=$NW"; MCMCB200_NUTS_PERSIST_NW=$NW timeout 300 python tools/prof_c4.py $C 200 200; done
done 2>&1 | tee gpurun_out/c4_$tag.log
echo "C=2048 NW=8"; MCMCB200_NUTS_PERSIST_NW=8 timeout 300 python tools/prof_c4.py 2048 200 200 2>&1 | tee -a gpurun_out/c4_$tag.log
echo "C=2048 NW=16"; MCMCB200_NUTS_PERSIST_NW=16 timeout 300 python tools/prof_c4.py 2048 200 200 2>&1 | tee -a gpurun_out/c4_$tag.log
