#!/bin/bash
# round 2, call 30: final state -- full GPU suite, smoke, ncu captures of the 7-pass step
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_s30_tests.log 2>&1
tail -9 gpurun_out/r02_s30_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s30_smoke.log 2>&1
tail -2 gpurun_out/r02_s30_smoke.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:table_gram_kernel3 -c 1 -o gpurun_out/r02_s30_k1 -f \
    python tools/profile_step.py 10000 1000000 0 > gpurun_out/r02_s30_ncu_k1.log 2>&1
tail -2 gpurun_out/r02_s30_ncu_k1.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_s30_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --no-eigen > gpurun_out/r02_s30_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r02_s30_bench_under_ncu.log | cut -c1-300
