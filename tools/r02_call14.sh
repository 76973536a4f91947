#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02_s14b_tests.log 2>&1
python bench.py > gpurun_out/r02_s14b_bench.json 2> gpurun_out/r02_s14b_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_s14b_ref.json 2> gpurun_out/r02_s14b_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_s14b_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --no-eigen > gpurun_out/r02_s14b_bench_under_ncu.log 2>&1
tail -8 gpurun_out/r02_s14b_tests.log; tail -3 gpurun_out/r02_s14b_bench.err; head -c 2500 gpurun_out/r02_s14b_bench.json; echo; cat gpurun_out/r02_s14b_ref.json | head -c 1500
