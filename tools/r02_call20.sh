#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pair_count_kernel -c 1 -o gpurun_out/r02_s20_ibs -f \
    python tools/quick_perf.py 16384 262144 ibs > gpurun_out/r02_s20_ibs.log 2>&1
tail -2 gpurun_out/r02_s20_ibs.log
