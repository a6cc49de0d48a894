#!/bin/bash
# session 26: full validation of the round's final kernels -- smoke, every -m gpu test, bench line (with cpu_baseline), reference arm
mkdir -p gpurun_out
run() { local n=$1 t=$2; shift 2; timeout $t "$@" > gpurun_out/$n.log 2>&1; echo "$n exit $? | $(tail -1 gpurun_out/$n.log | cut -c1-300)"; }
run smoke 240 python __graft_entry__.py smoke
grep -q "smoke ok" gpurun_out/smoke.log || { tail -5 gpurun_out/smoke.log; exit 1; }
run test_gpu_all 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider
grep -E "FAILED|Error|assert" gpurun_out/test_gpu_all.log | head -12
timeout 900 python bench.py > gpurun_out/bench_s26.json 2> gpurun_out/bench_s26.err; echo "bench exit $?"
tail -3 gpurun_out/bench_s26.err
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_s26.json 2> gpurun_out/bench_ref_s26.err; echo "reference arm exit $?"
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_s26.json").read().strip().splitlines() if l.startswith("{")][-1])
print("events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4), "frac", round(j["roofline"]["frac"], 3),
      "traffic", j["roofline"]["traffic"], "vit ms", round(j["stages"]["vit_encode"]["ms"], 2), "prefill ms", round(j["stages"]["prefill"]["ms"], 2))
for k, v in j.get("batched", {}).items():
    r = v.get("roofline", {})
    print(k, v["unit"], round(v["value"], 2), "ms", round(v["ms"], 1), "ms/step", round(r.get("ms_per_decode_step", 0), 3), "frac", round(r.get("frac", 0), 3))
print("cpu_baseline", {k: v for k, v in j.get("cpu_baseline", {}).items() if k != "sample"})
r = json.loads([l for l in open("gpurun_out/bench_ref_s26.json").read().strip().splitlines() if l.startswith("{")][-1])
print("reference arm", round(r["value"], 4), "spread", r["spread"], "ms_per_step", round(r["ms_per_step"]))
PY
