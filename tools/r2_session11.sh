#!/bin/bash
# Round-2 GPU session 11: prefill graphs -- parity suite, bench line, then the ncu evidence of the same bench command.
mkdir -p gpurun_out
run() { local n=$1 t=$2; shift 2; timeout $t "$@" > gpurun_out/$n.log 2>&1; echo "$n exit $? | $(tail -1 gpurun_out/$n.log | cut -c1-300)"; }
run smoke 240 python __graft_entry__.py smoke
grep -q "smoke ok" gpurun_out/smoke.log || { tail -5 gpurun_out/smoke.log; exit 1; }
run test_gpu_all 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider
grep -E "FAILED|Error|assert" gpurun_out/test_gpu_all.log | head -12
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench_s11.json 2> gpurun_out/bench_s11.err; echo "bench exit $?"
tail -3 gpurun_out/bench_s11.err
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_s11.json").read().strip().splitlines() if l.startswith("{")][-1])
print("events/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "ms/decode", round(j["roofline"]["ms_per_decode_step"], 4), "frac", round(j["roofline"]["frac"], 3),
      "vit ms", round(j["stages"]["vit_encode"]["ms"], 2), "prefill ms", round(j["stages"]["prefill"]["ms"], 2))
for k, v in j.get("batched", {}).items():
    r = v.get("roofline", {})
    print(k, v["unit"], round(v["value"], 2), "ms", round(v["ms"], 1), "ms/step", round(r.get("ms_per_decode_step", 0), 3), "frac", round(r.get("frac", 0), 3))
PY
export TRACE_B200_WATCHDOG_MS=120000
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3200 --csv \
  --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-batched > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $? rows $(wc -l < gpurun_out/r02_launches.csv)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decode_stream_kernel -s 40 -c 1 -f -o gpurun_out/r02_decode_step \
  python tools/profile_step.py --frames 128 --max-new 48 > gpurun_out/ncu_decode.log 2>&1
echo "decode capture exit $?"
ncu -i gpurun_out/r02_decode_step.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_decode_step.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_attn python tools/profile_step.py --frames 32 --max-new 2 > gpurun_out/ncu_attn.log 2>&1
ncu -i gpurun_out/r02_attn.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_attn_tc.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc2_kernel -s 8 -c 4 -f -o gpurun_out/r02_gemm python tools/profile_step.py --frames 32 --max-new 2 > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/r02_gemm.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_gemm_tc2.csv 2>/dev/null
ls -la gpurun_out/*.csv | head
rm -f gpurun_out/r02_attn.ncu-rep gpurun_out/r02_gemm.ncu-rep
