#!/bin/bash
# Final visit of the round: full parity suite, the bench line, the ncu launch list of the bench command, kernel summaries.
tag=${1:-r2final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$tag.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs "" > gpurun_out/bench_under_ncu_$tag.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$tag.csv
bash tools/gpu_prof2.sh $tag
