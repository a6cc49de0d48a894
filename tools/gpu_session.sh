#!/bin/bash
# One gpurun call: parity suite, decode A/B over the env switches of the fused decode step, bench line.
# Every leg has its own timeout so that a hung kernel cannot eat the whole box lease.
mkdir -p gpurun_out
bash tests/run_gpu_suite.sh
ab() {  # name, env...
  local name=$1; shift
  ( env "$@" timeout 240 python tools/profile_step.py --frames 128 --max-new 96 --warm 1 > gpurun_out/ab_$name.log 2>&1; echo "exit $?" >> gpurun_out/ab_$name.log )
  echo "== $name: $(grep DECODE_TIMING gpurun_out/ab_$name.log | tail -1) $(tail -1 gpurun_out/ab_$name.log)"
}
ab relay0 TRACE_B200_TM_RELAY=0
ab relay1_gate2 TRACE_B200_TM_RELAY=1 TRACE_B200_GATE_MODE=2
ab relay1_gate1 TRACE_B200_TM_RELAY=1 TRACE_B200_GATE_MODE=1
ab relay1_gate0 TRACE_B200_TM_RELAY=1 TRACE_B200_GATE_MODE=0
ab relay0_tl TRACE_B200_TM_RELAY=0 TRACE_B200_MEGA_TIMING=1
ab relay1_tl TRACE_B200_TM_RELAY=1 TRACE_B200_MEGA_TIMING=1
grep -h "cta\|qkv:\|o:\|gateup:\|down:" gpurun_out/ab_relay0_tl.log gpurun_out/ab_relay1_tl.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json
