#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
for i in 1 2 3 4 5; do timeout 300 python tools/prof_c5.py 2 6; done 2>&1 | tee gpurun_out/c5_times.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
