#!/bin/bash
# round 2, call 35: the state the round ends with -- full GPU suite and smoke
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r02_s35_tests.log 2>&1
tail -4 gpurun_out/r02_s35_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
