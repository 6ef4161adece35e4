#!/bin/bash
tag=${1:-r2}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none --import-source on"
cap() {
    local name=$1 rx=$2 skip=$3 per=$4; shift 4
    $NCU -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/${tag}_$name.log 2>&1
    tail -1 gpurun_out/${tag}_$name.log
    python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep $per > gpurun_out/${tag}_$name.ncu_summary.txt 2>&1
    rm -f /tmp/ncu/$name.ncu-rep
}
cap nuts_c4 nuts_kernel 0 6774000 python tools/prof_c4.py 1184 20 20
cap rmhmc_c5 rmhmc_cta_kernel 0 6144 python tools/prof_c5.py 1 2
