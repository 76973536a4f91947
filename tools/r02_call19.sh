#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_s19_bench8.json 2> gpurun_out/r02_s19_bench8.err
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02_s19_bench8.err | tail -5
python - <<'PY'
import json
for f in ("bench8",):
    try:
        line = [l for l in open(f"gpurun_out/r02_s19_{f}.json") if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_parts"], "parity", d["parity"]["max_rel_err"], d["roofline"]["fixed_point"])
        for k, v in d["extra"].items():
            print("  ", k, json.dumps(v)[:900])
    except Exception as e:
        print(f, "parse error", e)
PY
