#!/bin/bash
# Round-2 GPU session 1: first run of the streamed decode step (decode_stream.cu) -- smoke, parity suites incl. the
# full-width fixtures, the round-1 kernel on the same fixtures (A/B), and a quick bench line of both.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt
run() { # name timeout cmd...
  local n=$1 t=$2; shift 2
  timeout $t "$@" > gpurun_out/$n.log 2>&1
  echo "$n exit $? | $(tail -1 gpurun_out/$n.log | cut -c1-300)"
}
run smoke_v2 240 python __graft_entry__.py smoke
if ! grep -q "smoke ok" gpurun_out/smoke_v2.log; then
  tail -5 gpurun_out/smoke_v2.log
  export TRACE_B200_COOP=0
  run smoke_v2_nocoop 240 python __graft_entry__.py smoke
  if ! grep -q "smoke ok" gpurun_out/smoke_v2_nocoop.log; then
    tail -5 gpurun_out/smoke_v2_nocoop.log
    unset TRACE_B200_COOP
    TRACE_B200_DECODE=1 run smoke_v1 240 python __graft_entry__.py smoke
  fi
fi
for f in test_gpu_e2e test_gpu_wide test_gpu_serving; do
  run $f 900 python -m pytest tests/$f.py -m gpu -q -x --timeout 400 -p no:cacheprovider -s
  grep -E "FAILED|Error|assert|worst|graph loop|watchdog" gpurun_out/$f.log | head -12
done
TRACE_B200_DECODE=1 run test_gpu_wide_v1 900 python -m pytest tests/test_gpu_wide.py -m gpu -q --timeout 400 -p no:cacheprovider -s -k wide_case
grep -E "FAILED|Error|assert|worst|graph loop" gpurun_out/test_gpu_wide_v1.log | head -12
run bench_v2 400 python bench.py --steps 1 --warmup 1 --skip-cpu-baseline
TRACE_B200_DECODE=1 run bench_v1 400 python bench.py --steps 1 --warmup 1 --skip-cpu-baseline
python - <<'PY'
import json
for n in ("bench_v2", "bench_v1"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{n}.log").read().strip().splitlines() if l.startswith("{")][-1])
        print(n, "events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4),
              "frac", round(j["roofline"]["frac"], 3), "vit ms", round(j["stages"]["vit_encode"]["ms"], 2), "prefill ms", round(j["stages"]["prefill"]["ms"], 2))
    except Exception as e:
        print(n, "unreadable", e)
PY
