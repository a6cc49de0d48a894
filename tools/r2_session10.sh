#!/bin/bash
# Round-2 GPU session 10: dense k-outer with up-front flags -- parity, batch 4 / 8 timelines, full bench line + reference arm.
mkdir -p gpurun_out
run() { local n=$1 t=$2; shift 2; timeout $t "$@" > gpurun_out/$n.log 2>&1; echo "$n exit $? | $(tail -1 gpurun_out/$n.log | cut -c1-300)"; }
run smoke 240 python __graft_entry__.py smoke
grep -q "smoke ok" gpurun_out/smoke.log || { tail -5 gpurun_out/smoke.log; exit 1; }
run test_gpu 1200 python -m pytest tests/test_gpu_wide.py tests/test_gpu_serving.py -m gpu -q --timeout 400 -p no:cacheprovider
grep -E "FAILED|Error|assert" gpurun_out/test_gpu.log | head -12
export TRACE_B200_MEGA_TIMING=1 MAXNEW=96
BATCH=4 bash tools/ab.sh b4:
BATCH=8 bash tools/ab.sh b8:
unset TRACE_B200_MEGA_TIMING
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
tail -3 gpurun_out/bench_full.err
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm exit $?"
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_full.json").read().strip().splitlines() if l.startswith("{")][-1])
print("events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4), "frac", round(j["roofline"]["frac"], 3),
      "vit ms", round(j["stages"]["vit_encode"]["ms"], 2), "prefill ms", round(j["stages"]["prefill"]["ms"], 2))
for k, v in j.get("batched", {}).items():
    r = v.get("roofline", {})
    print(k, v["unit"], round(v["value"], 2), "ms", round(v["ms"], 1), "ms/step", round(r.get("ms_per_decode_step", 0), 3), "frac", round(r.get("frac", 0), 3))
print("cpu_baseline", {k: v for k, v in j.get("cpu_baseline", {}).items() if k != "sample"})
r = json.loads([l for l in open("gpurun_out/bench_ref.json").read().strip().splitlines() if l.startswith("{")][-1])
print("reference arm", round(r["value"], 4), "spread", r["spread"], "ms_per_step", round(r["ms_per_step"]), r["cpu_baseline"]["sample"][:400])
PY
