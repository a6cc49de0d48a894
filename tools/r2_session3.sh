#!/bin/bash
# Round-2 GPU session 3: O^T attention + setmaxnreg build -- parity first, then timeline at batch 1 / 4 / 8.
mkdir -p gpurun_out
run() { local n=$1 t=$2; shift 2; timeout $t "$@" > gpurun_out/$n.log 2>&1; echo "$n exit $? | $(tail -1 gpurun_out/$n.log | cut -c1-300)"; }
run smoke 240 python __graft_entry__.py smoke
grep -q "smoke ok" gpurun_out/smoke.log || { tail -5 gpurun_out/smoke.log; exit 1; }
run test_gpu_wide 900 python -m pytest tests/test_gpu_wide.py -m gpu -q -x --timeout 400 -p no:cacheprovider -s
grep -E "FAILED|Error|assert|worst|graph loop|watchdog" gpurun_out/test_gpu_wide.log | head -12
run test_gpu_e2e 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout 400 -p no:cacheprovider
export TRACE_B200_MEGA_TIMING=1 MAXNEW=64
bash tools/ab.sh s5:
BATCH=4 bash tools/ab.sh b4:
BATCH=8 bash tools/ab.sh b8:
