#!/bin/bash
# One GPU visit: parity suite, then the bench line.  Output under gpurun_out/ (merged back by gpurun).
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -L; nproc
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
