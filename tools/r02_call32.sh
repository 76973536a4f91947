#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_rounding.py -m gpu -q -x -k "plan_statistics" > gpurun_out/r02_s32_tests.log 2>&1
tail -12 gpurun_out/r02_s32_tests.log
timeout 200 python tools/quick_perf.py 10000 1000000 gcta,eigmix 0.005 3 > gpurun_out/r02_s32_quick.log 2>&1
cat gpurun_out/r02_s32_quick.log
