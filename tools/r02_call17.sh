#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_s17_bench8.json 2> gpurun_out/r02_s17_bench8.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$TR4 --master-port 29552 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02_s17_bench4.json 2> gpurun_out/r02_s17_bench4.err
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02_s17_bench8.err | tail -5; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02_s17_bench4.err | tail -5
python - <<'PY'
import json
for f in ("bench8", "bench4"):
    try:
        line = [l for l in open(f"gpurun_out/r02_s17_{f}.json") if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_parts"], "parity", d["parity"]["max_rel_err"], d["roofline"]["fixed_point"], d["clocks"])
        for k, v in d["extra"].items():
            print("  ", k, json.dumps(v)[:700])
    except Exception as e:
        print(f, "parse error", e)
PY
