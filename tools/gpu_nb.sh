#!/bin/bash
tag=${1:-nb1}
mkdir -p gpurun_out
MCMCB200_DEBUG=1 timeout 300 python tools/prof_c4.py 2368 100 100 2>&1 | tee gpurun_out/c4_$tag.log
