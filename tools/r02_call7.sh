#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_gpu_rshim.py -m gpu -q -x --durations=8 > gpurun_out/r02_s7_tests.log 2>&1
tail -30 gpurun_out/r02_s7_tests.log
