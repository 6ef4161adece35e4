#!/bin/bash
tag=${1:-nb6}
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_ls_step -s 12000 -c 1 -f -o /tmp/ncu/ls python tools/prof_c4.py 4096 60 60 > gpurun_out/${tag}_ls.log 2>&1
tail -2 gpurun_out/${tag}_ls.log
python tools/ncu_lines.py /tmp/ncu/ls.ncu-rep 70 > gpurun_out/${tag}_ls.lines.txt 2>&1
ncu -i /tmp/ncu/ls.ncu-rep --page source --csv --print-source cuda 2>/dev/null | head -5 > gpurun_out/${tag}_ls.srchead.txt
