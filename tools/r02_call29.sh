#!/bin/bash
# round 2, call 29: full GPU suite, smoke, the default bench line and a config-5 slice (rank 0 of 8) under the
# 'auto' rounding mode
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_s29_tests.log 2>&1
tail -14 gpurun_out/r02_s29_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s29_smoke.log 2>&1
tail -2 gpurun_out/r02_s29_smoke.log
timeout 400 python bench.py > gpurun_out/r02_s29_bench.json 2> gpurun_out/r02_s29_bench.err
tail -3 gpurun_out/r02_s29_bench.err
python - <<'PY'
import json
line = [l for l in open("gpurun_out/r02_s29_bench.json") if l.startswith("{")][-1]
d = json.loads(line)
print("step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "\nroofline", d["roofline"], "\nclocks", d["clocks"], "\ncpu", d["cpu_baseline"], "\neigen", d["eigen_top32_ms"], "\nparity", d["parity"]["max_rel_err"], d["parity"]["ok"])
PY
timeout 240 python tools/c5_tiled.py --samples 500000 --snps 800000 --world 8 --rank 0 --max-windows 2 > gpurun_out/r02_s29_c5_slice.json 2> gpurun_out/r02_s29_c5.err
tail -2 gpurun_out/r02_s29_c5.err; tail -1 gpurun_out/r02_s29_c5_slice.json
