#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/r02_s11_tests.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29521 tools/config_run.py --est king --samples 20000 --snps 100000 --rows 4096 --engine tensor > gpurun_out/r02_s11_king_tensor.json 2> gpurun_out/r02_s11_a.err
$TR --master-port 29522 tools/config_run.py --est king --samples 20000 --snps 100000 --rows 4096 --engine bits > gpurun_out/r02_s11_king_bits.json 2> gpurun_out/r02_s11_b.err
$TR --master-port 29523 tools/config_run.py --est ibs --samples 20000 --snps 100000 --reduce nccl > gpurun_out/r02_s11_ibs_nccl.json 2> gpurun_out/r02_s11_c.err
$TR --master-port 29524 tools/config_run.py --est ibs --samples 20000 --snps 100000 > gpurun_out/r02_s11_ibs_peer.json 2> gpurun_out/r02_s11_d.err
$TR --master-port 29525 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_s11_bench2.json 2> gpurun_out/r02_s11_bench2.err
SNPREL_REDUCE=nccl $TR --master-port 29526 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/r02_s11_bench2_nccl.json 2> gpurun_out/r02_s11_bench2_nccl.err
$TR --master-port 29527 tools/c5_tiled.py --samples 40000 --snps 100000 --rows 2048 > gpurun_out/r02_s11_c5small.json 2> gpurun_out/r02_s11_e.err
tail -5 gpurun_out/r02_s11_tests.log
for f in king_tensor king_bits ibs_nccl ibs_peer c5small; do echo "== $f"; cat gpurun_out/r02_s11_$f.json; done
for f in a b c d e bench2 bench2_nccl; do tail -3 gpurun_out/r02_s11_$f.err | grep -v "OMP_NUM\|\*\*\*\*" ; done
python - <<'PY'
import json
for f in ("bench2", "bench2_nccl"):
    try:
        d = json.load(open(f"gpurun_out/r02_s11_{f}.json"))
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("extra"))
    except Exception as e:
        print(f, "parse error", e)
PY
