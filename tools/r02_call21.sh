#!/bin/bash
mkdir -p gpurun_out
./build/peaks_r02 > gpurun_out/r02_s21_peaks.json 2> gpurun_out/r02_s21_peaks.err
cp gpurun_out/r02_s21_peaks.json profiles/peaks_r02.json
python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02_s21_tests.log 2>&1
python bench.py > gpurun_out/r02_s21_bench.json 2> gpurun_out/r02_s21_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s21_smoke.log 2>&1
cat gpurun_out/r02_s21_peaks.json; tail -8 gpurun_out/r02_s21_tests.log; tail -2 gpurun_out/r02_s21_smoke.log; tail -3 gpurun_out/r02_s21_bench.err
python - <<'PY'
import json
line = [l for l in open("gpurun_out/r02_s21_bench.json") if l.startswith("{")][-1]
d = json.loads(line)
print("step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "\nroofline", d["roofline"], "\nclocks", d["clocks"], "\ncpu", d["cpu_baseline"], "\neigen", d["eigen_top32_ms"])
print(json.dumps(d["extra"])[:2500])
PY
