#!/bin/bash
# session 38: final kernels -- memcheck / synccheck again (attention changed: P in TMEM), full -m gpu suite
mkdir -p gpurun_out
export TRACE_B200_WATCHDOG_MS=600000
for tool in memcheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "sanitize $tool smoke exit $? | $(grep -E 'ERROR SUMMARY|smoke ok' gpurun_out/sanitize_${tool}_smoke.log | tr '\n' ' ' | cut -c1-300)"
done
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "gemm or attention or slot or norm" > gpurun_out/sanitize_memcheck_ops.log 2>&1
echo "sanitize memcheck ops exit $? | $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_memcheck_ops.log | tr '\n' ' ' | cut -c1-300)"
unset TRACE_B200_WATCHDOG_MS
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "test_gpu_all exit $? | $(tail -1 gpurun_out/test_gpu_all.log)"
