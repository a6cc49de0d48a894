#!/bin/bash
# ncu evidence for one generate() of the bench workload (6 new tokens): launch list with DRAM bytes, and --set full
# captures of the decode step kernel and of the tcgen05 GEMM (tensor-pipe utilisation of the ViT / prefill tiles).
mkdir -p gpurun_out
CMD="python tools/profile_step.py --frames 128 --max-new 6"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_step -s 2 -c 1 -f -o gpurun_out/prof_decode $CMD > gpurun_out/ncu_decode.log 2>&1; echo "decode full exit $?"
timeout 400 ncu --set full --clock-control none -k regex:gemm_tc -s 60 -c 6 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1; echo "gemm full exit $?"
timeout 300 ncu --set full --clock-control none -k regex:attn_tc -s 10 -c 2 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1; echo "attn full exit $?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
