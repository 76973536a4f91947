#!/bin/bash
mkdir -p gpurun_out
timeout 95 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zfull_size.py -m gpu -q -x -k "not long_k and not streamed and not top32 and not ibs_counts_full" > gpurun_out/r02_s37_tests.log 2>&1
tail -3 gpurun_out/r02_s37_tests.log
