#!/bin/bash
# Runs the GPU parity suite one file per process (a faulting kernel cannot poison later files) with hard timeouts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_serving.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit $?" | tee -a gpurun_out/summary.txt
  tail -40 gpurun_out/$n.log
done
