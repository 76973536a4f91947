#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r02_s10_tests.log 2>&1
python bench.py > gpurun_out/r02_s10_bench.json 2> gpurun_out/r02_s10_bench.err
timeout 600 python tools/k1_variants.py 10000 1000000 0,256,16384,49152,0 > gpurun_out/r02_s10_variants.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:table_gram_kernel3 -c 1 -o gpurun_out/r02_s10_k1 -f \
    python tools/profile_step.py 10000 1000000 0 > gpurun_out/r02_s10_ncu_k1.log 2>&1
tail -8 gpurun_out/r02_s10_tests.log; cat gpurun_out/r02_s10_variants.log; tail -3 gpurun_out/r02_s10_bench.err; head -c 1500 gpurun_out/r02_s10_bench.json
