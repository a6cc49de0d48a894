#!/bin/bash
# session 39: ncu --set full of one batched decode step (4 and 8 rows) for the DRAM traffic of the batched roofline records
mkdir -p gpurun_out
export TRACE_B200_WATCHDOG_MS=120000
for b in 8 4; do
  timeout 600 ncu --set full --clock-control none -k regex:decode_stream_kernel -s 20 -c 1 -f -o gpurun_out/r02_decode_step_b$b \
    python tools/profile_step.py --frames 128 --max-new 28 --batch $b > gpurun_out/ncu_decode_b$b.log 2>&1
  echo "capture b$b exit $?"
  ncu -i gpurun_out/r02_decode_step_b$b.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_decode_step_b$b.csv 2>/dev/null
  rm -f gpurun_out/r02_decode_step_b$b.ncu-rep
done
python tools/ncu_pick.py gpurun_out/r02_ncu_full_decode_step_b8.csv gpu__time_duration.sum "dram__bytes_read.sum " "dram__bytes_write.sum " | cut -c1-200
python tools/ncu_pick.py gpurun_out/r02_ncu_full_decode_step_b4.csv gpu__time_duration.sum "dram__bytes_read.sum " "dram__bytes_write.sum " | cut -c1-200
