#!/bin/bash
# round 2, call 28 (2 GPUs): SNP origins of the sharded paths under the 'auto' rounding mode
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02_s28_tests.log 2>&1
tail -3 gpurun_out/r02_s28_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_s28_bench2.json 2> gpurun_out/r02_s28_bench2.err
tail -3 gpurun_out/r02_s28_bench2.err
python - <<'PY'
import json
line = [l for l in open("gpurun_out/r02_s28_bench2.json") if l.startswith("{")][-1]
d = json.loads(line)
print("bench2 step", d["ms_per_step"], d["value"], d["roofline"]["fixed_point"], d["e2e"]["ms_per_step"], d["extra"].get("strong", {}).get("ms_per_step"), d["parity"])
PY
