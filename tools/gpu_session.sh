#!/bin/bash
# One gpurun call: smoke -> parity suite -> kernel micro-benchmarks -> bench line (+ tcgen05-attention variant).
# Every leg has its own timeout; a failed smoke stops the session (a hung kernel must not eat the box lease).
mkdir -p gpurun_out
timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1 || exit 1
for f in tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_serving.py; do
  n=$(basename $f .py)
  timeout 400 python -m pytest $f -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit $? $(tail -1 gpurun_out/$n.log)"
  grep -E "FAILED|Error|assert" gpurun_out/$n.log | head -8
done
timeout 200 python tools/bench_kernels.py > gpurun_out/kernel_microbench.txt 2>&1; tail -14 gpurun_out/kernel_microbench.txt
timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json
TRACE_B200_ATTN_TC=0 timeout 300 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_n1_attn_mma.json 2> gpurun_out/bench_n1_attn_mma.err; echo "bench(mma.sync attention) exit $?"
python - <<'PY'
import json
for n in ("bench_n1", "bench_n1_attn_mma"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4),
              "frac", round(j["roofline"]["frac"], 3), "vit", {k: round(v, 3) for k, v in j["stages"]["vit_encode"].items() if isinstance(v, float)},
              "prefill ms", round(j["stages"]["prefill"]["ms"], 2))
    except Exception as e:
        print(n, "unreadable", e)
PY
