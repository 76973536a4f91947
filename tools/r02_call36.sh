#!/bin/bash
# round 2, call 36: auxiliary byte tables of sample_stats_kernel summed three rows per byte lane
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_rounding.py tests/test_gpu_shard_algebra.py -m gpu -q -x > gpurun_out/r02_s36_tests.log 2>&1
tail -3 gpurun_out/r02_s36_tests.log
timeout 60 python tools/quick_perf.py 10000 1000000 pca 0.005 4 > gpurun_out/r02_s36_quick.log 2>&1
cat gpurun_out/r02_s36_quick.log
