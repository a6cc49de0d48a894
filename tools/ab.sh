#!/bin/bash
# usage: tools/ab.sh name:ENV=V,ENV2=V ...   -- one profile_step run (96 decode steps, B=1, 128 frames) per variant
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  ( env $(echo $envs | tr ',' ' ') timeout 240 python tools/profile_step.py --frames 128 --max-new ${MAXNEW:-96} --warm 1 --batch ${BATCH:-1} > gpurun_out/ab_$name.log 2>&1; echo "exit $?" >> gpurun_out/ab_$name.log )
  echo "== $name: $(grep DECODE_TIMING gpurun_out/ab_$name.log | tail -1) $(tail -1 gpurun_out/ab_$name.log)"
  grep -h "cta\|qkv:\|o:\|gateup:\|down:" gpurun_out/ab_$name.log
done
