#!/bin/bash
# 2-GPU call: diag-bound headroom + cached IPC exchange
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard_algebra.py tests/test_gpu_dist.py tests/test_gpu_multi.py -m gpu -q -x --durations=5 -k "not r_entry" > gpurun_out/r02_s16_tests.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_s16_bench2.json 2> gpurun_out/r02_s16_bench2.err
$TR --master-port 29542 tools/config_run.py --est king --samples 40000 --snps 200000 --rows 8192 --engine tensor > gpurun_out/r02_s16_king_tensor.json 2> gpurun_out/r02_s16_a.err
$TR --master-port 29543 tools/config_run.py --est king --samples 40000 --snps 200000 --rows 8192 --engine tensor --reduce nccl > gpurun_out/r02_s16_king_tensor_nccl.json 2> gpurun_out/r02_s16_b.err
python - > gpurun_out/r02_s16_plan8m.log 2>&1 <<'PY'
# the format the library picks for 10k x 8M SNPs (what 8 weak-scaled ranks agree on): plan statistics of 1M SNPs x 8
import snprelate_b200 as S
c = S.Context(0); c.geno_begin(10000, 1000000); c.geno_synth(1000000)
p = c.plan_local(0)
print("1M : sum_bound %.4g diag_bound %.4g sum_rest %.4g err_weight %.4g" % (p.sum_bound, p.diag_bound, p.sum_rest, p.err_weight))
c.accumulate(0, p); q = c.last_plan(); print("1M format: digits", q.digits, q.digits_w, "frac", q.frac_bits, q.frac_bits_w)
for k in ("sum_bound", "err_weight", "scale", "diag_bound", "sum_rest"):
    setattr(p, k, 8 * getattr(p, k))
p.total_missing *= 8; p.max_missing *= 8; p.n_snp *= 8
c.accumulate(0, p); q = c.last_plan(); print("8M format: digits", q.digits, q.digits_w, "frac", q.frac_bits, q.frac_bits_w)
PY
tail -5 gpurun_out/r02_s16_tests.log; cat gpurun_out/r02_s16_plan8m.log
for f in king_tensor king_tensor_nccl; do tail -1 gpurun_out/r02_s16_$f.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['job_s'], d['phases_s_this_rank'], d['parity'])"; done
python - <<'PY'
import json
line = [l for l in open("gpurun_out/r02_s16_bench2.json") if l.startswith("{")][-1]
d = json.loads(line)
print("bench2 step", d["ms_per_step"], d["roofline"]["fixed_point"], d["extra"]["reduction"], d["extra"]["strong"]["ms_per_step"], d["parity"]["max_rel_err"])
PY
