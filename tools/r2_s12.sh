#!/bin/bash
# session 12: attention kernel with two threads per row -- op tests, micro-benchmark, wide parity, bench stages
mkdir -p gpurun_out
run() { local n=$1 t=$2; shift 2; timeout $t "$@" > gpurun_out/$n.log 2>&1; echo "$n exit $? | $(tail -1 gpurun_out/$n.log | cut -c1-300)"; }
run test_ops 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider
grep -E "FAILED|Error|assert" gpurun_out/test_ops.log | head -12
run attn_bench 300 python tools/attn_bench.py
cat gpurun_out/attn_bench.log
run test_wide 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_e2e.py -m gpu -q --timeout 400 -p no:cacheprovider
grep -E "FAILED|Error|assert" gpurun_out/test_wide.log | head -12
timeout 600 python bench.py --skip-cpu-baseline --skip-batched > gpurun_out/bench_s12.json 2> gpurun_out/bench_s12.err; echo "bench exit $?"
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_s12.json").read().strip().splitlines() if l.startswith("{")][-1])
print("events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4), "frac", round(j["roofline"]["frac"], 3),
      "vit ms", round(j["stages"]["vit_encode"]["ms"], 2), "prefill ms", round(j["stages"]["prefill"]["ms"], 2))
PY
