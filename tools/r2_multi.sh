#!/bin/bash
# 2-GPU session: NCCL tests + the bench at N=2 (launched as the driver launches it) + the reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/test_gpu_multi.log 2>&1; echo "test_gpu_multi exit $? | $(tail -1 gpurun_out/test_gpu_multi.log)"
grep -E "FAILED|Error|assert" gpurun_out/test_gpu_multi.log | head
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N=2 events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "e2e ms/step", round(j["e2e"]["ms_per_step"], 1), "h2d/step", j["e2e"]["h2d_bytes_per_step"],
      "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4), "frac", round(j["roofline"]["frac"], 3))
for k, v in j.get("batched", {}).items():
    print(k, "clips", v["clips"], "events/s", round(v["value"], 2), "ms", round(v["ms"], 1), "ms/step", round(v["roofline"]["ms_per_decode_step"], 3))
PY
