#!/bin/bash
# round 2, call 27: the 'auto' rounding mode (randomised rounding where it saves a tensor pass)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_rounding.py tests/test_gpu_shard_algebra.py tests/test_gpu_parity.py -x -q -m gpu \
    -k "rounding or shard or streamed or row_windows or hapmap_config1" > gpurun_out/r02_s27_tests.log 2>&1
tail -5 gpurun_out/r02_s27_tests.log
timeout 200 python tools/rounding_check.py 10000 1000000 nearest,random,auto 0,1000003 > gpurun_out/r02_s27_rounding.log 2>&1
cat gpurun_out/r02_s27_rounding.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_s27_bench.json 2> gpurun_out/r02_s27_bench.err
tail -3 gpurun_out/r02_s27_bench.err
python - <<'PY'
import json
line = [l for l in open("gpurun_out/r02_s27_bench.json") if l.startswith("{")][-1]
d = json.loads(line)
print("step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["tensor_passes_per_step"], d["e2e"]["streamed_steps_fallbacks"], "\nroofline", d["roofline"], "\nclocks", d["clocks"], "\nparity", d["parity"]["max_rel_err"])
PY
