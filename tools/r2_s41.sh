#!/bin/bash
# session 41: vectorised prefill RoPE / KV write, 2 rows per CTA in the small-M norm -- parity + stage times
mkdir -p gpurun_out
run() { local n=$1 t=$2; shift 2; timeout $t "$@" > gpurun_out/$n.log 2>&1; echo "$n exit $? | $(tail -1 gpurun_out/$n.log | cut -c1-300)"; }
run test_gpu_all 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider
grep -E "FAILED|Error|assert" gpurun_out/test_gpu_all.log | head -12
timeout 600 python bench.py --skip-cpu-baseline --skip-batched > gpurun_out/bench_s41.json 2> gpurun_out/bench_s41.err; echo "bench exit $?"
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_s41.json").read().strip().splitlines() if l.startswith("{")][-1])
print("events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4), "frac", round(j["roofline"]["frac"], 3),
      "vit ms", round(j["stages"]["vit_encode"]["ms"], 2), "prefill ms", round(j["stages"]["prefill"]["ms"], 2))
PY
