#!/bin/bash
# session 18: ncu launch list of the bench command (own kernels only), compute-sanitizer targets
mkdir -p gpurun_out
export TRACE_B200_WATCHDOG_MS=120000
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:^(attn_tc|decode_stream|gemm_tc|norm_|resample_|rope_kv|slot_|splice|vit_)' -c 3200 --csv \
  --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-batched > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $? rows $(wc -l < gpurun_out/r02_launches.csv)"
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
bash tools/sanitize.sh
