#!/bin/bash
# Round-2 GPU session 2: in-kernel timeline of the streamed decode step and a few switches, batch 1 / 4 / 8.
export TRACE_B200_MEGA_TIMING=1 MAXNEW=64
bash tools/ab.sh s5: s4:TRACE_B200_STAGES=4 nocoop:TRACE_B200_COOP=0 sleep200:TRACE_B200_POLL_SLEEP=200
BATCH=4 bash tools/ab.sh b4:
BATCH=8 bash tools/ab.sh b8:
