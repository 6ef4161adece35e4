#!/bin/bash
# ncu captures of the round's kernels (one GPU).  The reports are summarised ON THE BOX (tools/ncu_summary.py) and deleted:
# only the text summaries travel back (gpurun_out/ is capped at 64 MiB).
tag=${1:-r2}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, skip, per-unit divisor, command...
    local name=$1 rx=$2 skip=$3 per=$4; shift 4
    $NCU -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/${tag}_$name.log 2>&1
    tail -1 gpurun_out/${tag}_$name.log
    python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep $per > gpurun_out/${tag}_$name.ncu_summary.txt 2>&1
    rm -f /tmp/ncu/$name.ncu-rep
}
cap nuts_c4 nuts_kernel 0 6774000 python tools/prof_c4.py 1184 20 20
cap rmhmc_c5 rmhmc_cta_kernel 0 6144 python tools/prof_c5.py 1 2
cap hmc_c2_512 hmc_pipe_kernel 1 563200 python tools/prof_c2_strong.py 512
cap hmc_c2 hmc_pipe_kernel 1 4505600 python tools/prof_c2.py 2
ls -la gpurun_out | tail -12
