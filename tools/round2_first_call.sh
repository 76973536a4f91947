#!/bin/bash
# First GPU call of the next round: everything that was written at the end of round 1 without a GPU.
# Usage (from the repo root, inside gpurun):  bash tools/round2_first_call.sh
# Writes gpurun_out/r02_first_*.log ; ~4 minutes on one B200.
mkdir -p gpurun_out
{
  echo "== experimental paths (randomised rounding, two-stream launches, GDS bit-stream ingest)"
  SNPREL_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zz_experimental.py -m gpu -q 2>&1 | tail -15
  echo "== config-2 full-size test and the R-shim tests, with durations"
  python -m pytest tests/test_gpu_zfull_size.py tests/test_gpu_rshim.py -m gpu -q --durations=8 2>&1 | tail -20
  echo "== R-shim per-call timing"
  timeout 300 python tools/rshim_timing.py 2>&1 | tail -14
} > gpurun_out/r02_first_tests.log 2>&1
{
  echo "== pair kernels after the lop3 / IMAD rewrite (round 1: IBS 4.9e13, KING 3.4e13 pair-SNPs/s)"
  python tools/quick_perf.py 8192 262144 ibs,king,beta
  echo "== table Gram: default, two-stream launches (flag 8), randomised rounding"
  python - <<'PY'
import snprelate_b200 as S
c = S.Context(0)
c.geno_begin(10000, 1000000)
c.geno_synth(1000000, miss_rate=0.005)
for name, flags, rnd in (("default", 0, "nearest"), ("two streams", 8, "nearest"), ("random rounding", 0, "random"),
                         ("both", 8, "random")):
    c.debug_flags(flags)
    c.set_rounding(rnd)
    ms = min(c.time_accumulate(0, 1) for _ in range(3))
    p = c.last_plan()
    print(f"{name:16s} step {ms:7.1f} ms  hot {c.last_hot_kernel()[0]:7.1f} ms  digits U{p.digits} W{p.digits_w} D{p.digits_d}", flush=True)
PY
} > gpurun_out/r02_first_perf.log 2>&1
tail -40 gpurun_out/r02_first_tests.log gpurun_out/r02_first_perf.log
