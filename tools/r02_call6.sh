#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02_s6_tests.log 2>&1
python bench.py > gpurun_out/r02_s6_bench.json 2> gpurun_out/r02_s6_bench.err
tail -15 gpurun_out/r02_s6_tests.log; tail -5 gpurun_out/r02_s6_bench.err; cat gpurun_out/r02_s6_bench.json
