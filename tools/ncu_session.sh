#!/bin/bash
# ncu evidence for profiles/ (one GPU): (1) launch list of one generate() of the bench workload with DRAM bytes per launch,
# (2) --set full capture of one decode step (decode_stream_kernel<1>) and of one ViT GEMM / attention launch.
# Numbers printed by a run under ncu are never bench values; compare SHARES of the launch list, not absolute times.
mkdir -p gpurun_out
export TRACE_B200_WATCHDOG_MS=120000
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/r02_launches.csv python tools/profile_step.py --frames 128 --max-new 6 > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $? rows $(wc -l < gpurun_out/r02_launches.csv)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decode_stream_kernel -s 4 -c 1 -f -o gpurun_out/r02_decode_step \
  python tools/profile_step.py --frames 128 --max-new 8 > gpurun_out/ncu_decode.log 2>&1
echo "decode capture exit $?"
ncu -i gpurun_out/r02_decode_step.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_decode_step.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_attn python tools/profile_step.py --frames 32 --max-new 2 > gpurun_out/ncu_attn.log 2>&1
ncu -i gpurun_out/r02_attn.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_attn_tc.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc2_kernel -s 8 -c 4 -f -o gpurun_out/r02_gemm python tools/profile_step.py --frames 32 --max-new 2 > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/r02_gemm.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_gemm_tc2.csv 2>/dev/null
ls -la gpurun_out/*.csv | head
rm -f gpurun_out/r02_attn.ncu-rep gpurun_out/r02_gemm.ncu-rep     # keep the merge under the size limit; the decode report stays for source-level reading
