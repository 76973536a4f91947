#!/bin/bash
# 8-GPU call
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_s15_bench8.json 2> gpurun_out/r02_s15_bench8.err
$TR --master-port 29532 tools/c5_tiled.py > gpurun_out/r02_s15_c5.json 2> gpurun_out/r02_s15_c5.err
timeout 600 python tools/multi_perf.py > gpurun_out/r02_s15_multi.json 2> gpurun_out/r02_s15_multi.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "sharded or tiled or windows" > gpurun_out/r02_s15_tests.log 2>&1
tail -4 gpurun_out/r02_s15_tests.log; cat gpurun_out/r02_s15_multi.json; tail -3 gpurun_out/r02_s15_multi.err
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02_s15_c5.err | tail -3; cat gpurun_out/r02_s15_c5.json | tail -1 | cut -c1-1800
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02_s15_bench8.err | tail -5
python - <<'PY'
import json
try:
    line = [l for l in open("gpurun_out/r02_s15_bench8.json") if l.startswith("{")][-1]
    d = json.loads(line)
    print("bench8 step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["max_rel_err"])
    print(json.dumps(d["extra"])[:3500])
except Exception as e:
    print("bench8 parse error", e)
PY
