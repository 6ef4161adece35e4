#!/bin/bash
# Final visit of the round: full parity suite, the bench line, the ncu launch list of the bench command, kernel summaries.
tag=${1:-r2final}
mkdir -p gpurun_out /tmp/ncu
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$tag.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs "" > gpurun_out/bench_under_ncu_$tag.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$tag.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nuts_pc_kernel -c 1 -f -o /tmp/ncu/pc python tools/prof_c4.py 2368 15 15 > gpurun_out/${tag}_pc.log 2>&1
tail -2 gpurun_out/${tag}_pc.log
python tools/ncu_summary.py /tmp/ncu/pc.ncu-rep 10158720 > gpurun_out/${tag}_nuts_pc.ncu_summary.txt 2>&1
ncu -i /tmp/ncu/pc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for a,b,c in zip(h,u,v):
    if any(k in a for k in ('tensor_subpipe_dmma','lts__throughput.avg','lts__t_sector_hit_rate','sm__throughput.avg','l1tex__throughput.avg.pct_of_peak_sustained_elapsed')): print(a,b,c)
" >> gpurun_out/${tag}_nuts_pc.ncu_summary.txt 2>&1
