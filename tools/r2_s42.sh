#!/bin/bash
# session 42: refresh the ncu evidence on the round's final kernels -- launch list of the bench command, full captures of the ViT attention / GEMMs
mkdir -p gpurun_out
export TRACE_B200_WATCHDOG_MS=120000
timeout 900 ncu --set full --clock-control none -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_attn python tools/profile_step.py --frames 64 --max-new 2 > gpurun_out/ncu_attn.log 2>&1
ncu -i gpurun_out/r02_attn.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_attn_tc.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc2_kernel -s 8 -c 4 -f -o gpurun_out/r02_gemm python tools/profile_step.py --frames 64 --max-new 2 > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/r02_gemm.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_gemm_tc2.csv 2>/dev/null
rm -f gpurun_out/r02_attn.ncu-rep gpurun_out/r02_gemm.ncu-rep
python tools/ncu_pick.py gpurun_out/r02_ncu_full_attn_tc.csv gpu__time_duration.sum sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct launch__registers_per_thread | cut -c1-160
python tools/ncu_pick.py gpurun_out/r02_ncu_full_gemm_tc2.csv gpu__time_duration.sum sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active | cut -c1-160
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:^(attn_tc|decode_stream|gemm_tc|norm_|resample_|rope_kv|slot_|splice|vit_)' -c 2400 --csv \
  --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-batched > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $? rows $(wc -l < gpurun_out/r02_launches.csv)"
