#!/bin/bash
# round 2, call 31: GCTA / EIGMIX at config-2 size under the auto rounding mode + a sanity pass over the GPU suite
mkdir -p gpurun_out
timeout 200 python tools/quick_perf.py 10000 1000000 pca,gcta,eigmix > gpurun_out/r02_s31_quick.log 2>&1
cat gpurun_out/r02_s31_quick.log
timeout 300 python -m pytest tests -m gpu -q -x -k "not zfull and not long_k" > gpurun_out/r02_s31_tests.log 2>&1
tail -3 gpurun_out/r02_s31_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
