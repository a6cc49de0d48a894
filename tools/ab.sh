#!/bin/bash
# usage: tools/ab.sh name:ENV=V,ENV2=V ...   -- one profile_step run (96 decode steps, B=1, 128 frames) per variant
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader -lms 100 > gpurun_out/clk_$name.log 2>&1 &
  SMI=$!
  ( env AB_NAME=$name $(echo $envs | tr ',' ' ') timeout 240 python tools/profile_step.py --frames 128 --max-new ${MAXNEW:-96} --warm 1 --batch ${BATCH:-1} > gpurun_out/ab_$name.log 2>&1; echo "exit $?" >> gpurun_out/ab_$name.log )
  kill $SMI 2>/dev/null
  echo "   clocks(sm MHz, sorted tail): $(cut -d, -f1 gpurun_out/clk_$name.log | tr -d ' MHz' | sort -n | tail -12 | tr '\n' ' ') power: $(cut -d, -f3 gpurun_out/clk_$name.log | sort -n | tail -1)"
  echo "== $name: $(grep DECODE_TIMING gpurun_out/ab_$name.log | tail -1) $(tail -1 gpurun_out/ab_$name.log) $(grep -o "sha [0-9a-f]*" gpurun_out/ab_$name.log)"
  grep -h "cta\|per-CTA\|SM bandwidth" gpurun_out/ab_$name.log | grep -v Traceback
done
