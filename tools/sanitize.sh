#!/bin/bash
# compute-sanitizer targets (race / memory checking): the tiny-dims smoke (every kernel of the path incl. the
# cooperative decode step with its mbarrier rings and flag-in-data exchanges) under memcheck, racecheck and synccheck, and
# the tcgen05 GEMM / attention op tests under memcheck.  Logs -> gpurun_out/sanitize_*.log (summaries are copied to profiles/).
mkdir -p gpurun_out
export TRACE_B200_WATCHDOG_MS=600000
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "sanitize $tool smoke exit $? | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitize_${tool}_smoke.log | tr '\n' ' ' | cut -c1-300)"
done
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "gemm or attention or slot or norm" > gpurun_out/sanitize_memcheck_ops.log 2>&1
echo "sanitize memcheck ops exit $? | $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_memcheck_ops.log | tr '\n' ' ' | cut -c1-300)"
