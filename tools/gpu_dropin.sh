#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cpp_dropin.py tests/test_gpu_cabi.py -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dropin.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_dropin.log
