#!/bin/bash
# Round-2 GPU session 8: k-outer O / down projections (chunks multiplied as they arrive) -- parity + timeline + bench.
mkdir -p gpurun_out
run() { local n=$1 t=$2; shift 2; timeout $t "$@" > gpurun_out/$n.log 2>&1; echo "$n exit $? | $(tail -1 gpurun_out/$n.log | cut -c1-300)"; }
run smoke 240 python __graft_entry__.py smoke
grep -q "smoke ok" gpurun_out/smoke.log || { tail -5 gpurun_out/smoke.log; exit 1; }
run test_gpu_wide 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 400 -p no:cacheprovider
grep -E "FAILED|Error|assert" gpurun_out/test_gpu_wide.log | head -12
export TRACE_B200_MEGA_TIMING=1 MAXNEW=128
bash tools/ab.sh korder:
unset TRACE_B200_MEGA_TIMING
timeout 900 python bench.py --steps 2 --warmup 2 --skip-cpu-baseline --skip-batched > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; echo "bench exit $?"
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_b1.json").read().strip().splitlines() if l.startswith("{")][-1])
print("events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4), "frac", round(j["roofline"]["frac"], 3))
PY
