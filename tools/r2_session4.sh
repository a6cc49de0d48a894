#!/bin/bash
# Round-2 GPU session 4: per-CTA phase skew (two runs: is it tied to the SM?) and the full new bench line.
mkdir -p gpurun_out
export TRACE_B200_MEGA_TIMING=1 MAXNEW=48
bash tools/ab.sh skew1: skew2:
unset TRACE_B200_MEGA_TIMING
timeout 900 python bench.py --steps 2 --warmup 2 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_full.json").read().strip().splitlines() if l.startswith("{")][-1])
print("events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4), "frac", round(j["roofline"]["frac"], 3))
for k, v in j.get("batched", {}).items():
    print(k, "events/s", round(v["value"], 2), "ms", round(v["ms"], 1), "ms/step", round(v["roofline"]["ms_per_decode_step"], 3), "frac", round(v["roofline"]["frac"], 3))
print("cpu", j.get("cpu_baseline"))
PY
