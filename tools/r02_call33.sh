#!/bin/bash
# round 2, call 33: K1 at 7 passes -- per-item clock stamps by item class, and 4 / 6 / 8 / 12 SNP segments
mkdir -p gpurun_out
timeout 150 python tools/k1_trace.py 10000 1000000 > gpurun_out/r02_s33_k1_trace.log 2>&1
cat gpurun_out/r02_s33_k1_trace.log
timeout 200 python tools/k1_variants.py 10000 1000000 0,16384,24576,49152 > gpurun_out/r02_s33_variants.log 2>&1
cat gpurun_out/r02_s33_variants.log
