#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streamed or pca_golden or large_properties" --durations=5 > gpurun_out/r02_s12_tests.log 2>&1
python bench.py --no-extra --no-cpu --no-eigen > gpurun_out/r02_s12_bench.json 2> gpurun_out/r02_s12_bench.err
python bench.py --no-extra --no-cpu --no-eigen --sync-ingest > gpurun_out/r02_s12_bench_sync.json 2> gpurun_out/r02_s12_bench_sync.err
tail -30 gpurun_out/r02_s12_tests.log
python - <<'PY'
import json
for f in ("bench", "bench_sync"):
    try:
        line = [l for l in open(f"gpurun_out/r02_s12_{f}.json") if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_parts"], d["e2e"].get("streamed_steps_fallbacks"), d["parity"]["max_rel_err"], d["clocks"])
    except Exception as e:
        print(f, "parse error", e); print(open(f"gpurun_out/r02_s12_{f}.err").read()[-2000:])
PY
