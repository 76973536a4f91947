#!/bin/bash
# 2-GPU call: real peer access / NCCL paths
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_s8_topo.txt 2>&1
python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist.py tests/test_gpu_rshim.py -m gpu -q --durations=8 > gpurun_out/r02_s8_tests.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_s8_bench2.json 2> gpurun_out/r02_s8_bench2.err
python - > gpurun_out/r02_s8_multi_perf.log 2>&1 <<'PY'
import time, numpy as np
import snprelate_b200 as S
from snprelate_b200._lib import EST_IBS
# config 2 as one fixed problem on 2 devices of ONE process: accumulate + peer reduce + finish
for devs in ([0], [0, 1]):
    if len(devs) == 1:
        c = S.Context(0); c.geno_begin(10000, 1000000); c.geno_synth(1000000)
        t = time.perf_counter(); r = c.pca(genmat_only=True); dt = time.perf_counter() - t
        t = time.perf_counter(); r = c.pca(genmat_only=True); dt = time.perf_counter() - t
        print("1 device  pca genmat", round(dt * 1e3, 1), "ms"); c.close(); continue
    m = S.MultiContext(devs); m.geno_begin(10000, 1000000); m.geno_synth(1000000)
    for rep in range(3):
        t = time.perf_counter(); m.accumulate("Eigenstrat"); t1 = time.perf_counter(); r = m.ctx(0).pca(genmat_only=True); t2 = time.perf_counter()
        print(f"{len(devs)} devices accumulate+reduce {1e3*(t1-t):.1f} ms (reduce {m.last_reduce()[0]:.2f} ms, {m.last_reduce()[1]/1e9:.2f} GB over the links) finish {1e3*(t2-t1):.1f} ms")
    m.accumulate(EST_IBS); print("ibs reduce", m.last_reduce())
PY
tail -25 gpurun_out/r02_s8_tests.log; cat gpurun_out/r02_s8_multi_perf.log; tail -3 gpurun_out/r02_s8_bench2.err; cat gpurun_out/r02_s8_bench2.json | head -c 3000
