#!/bin/bash
mkdir -p gpurun_out
timeout 280 python tools/rounding_stress.py > gpurun_out/r02_s34_stress.log 2>&1
cat gpurun_out/r02_s34_stress.log | tail -14
