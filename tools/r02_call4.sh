#!/bin/bash
# round 2, call 4: full gpu suite on the committed state, bench, launch list, one full ncu capture of K1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02_s4_tests.log 2>&1
python bench.py > gpurun_out/r02_s4_bench.json 2> gpurun_out/r02_s4_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_s4_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r02_s4_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:table_gram_kernel3 -c 1 -o gpurun_out/r02_s4_k1 -f \
    python tools/profile_step.py 10000 1000000 0 > gpurun_out/r02_s4_ncu_k1.log 2>&1
tail -5 gpurun_out/r02_s4_tests.log; cat gpurun_out/r02_s4_bench.json
