#!/bin/bash
# Round-2 GPU session 7: is the per-CTA skew tied to the access pattern?  uniform partition vs closed-loop balance vs
# contiguous-stage experiment (garbage data, timing only), batch 1.
export TRACE_B200_MEGA_TIMING=1 MAXNEW=192
bash tools/ab.sh bal: nobal:TRACE_B200_BALANCE=0 packed:TRACE_B200_PACKED_EXPERIMENT=1,TRACE_B200_BALANCE=0
MAXNEW=100 BATCH=4 bash tools/ab.sh b4:
