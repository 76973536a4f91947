#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_s25_bench.json 2> gpurun_out/r02_s25_bench.err
ncu --set full --clock-control none --import-source on -k regex:table_gram_kernel3 -c 1 -o gpurun_out/r02_s25_k1 -f \
    python tools/profile_step.py 10000 1000000 0 > gpurun_out/r02_s25_ncu_k1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_s25_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --no-eigen > gpurun_out/r02_s25_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02_s25_bench.err
python - <<'PY'
import json
line = [l for l in open("gpurun_out/r02_s25_bench.json") if l.startswith("{")][-1]
d = json.loads(line)
print("step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["tensor_passes_per_step"], d["e2e"]["streamed_steps_fallbacks"], "\nroofline", d["roofline"], "\nclocks", d["clocks"], "\nparity", d["parity"]["max_rel_err"])
PY
