#!/bin/bash
# usage: tools/ncu_fb.sh <variant> <kernel regex> <tag> [rows]   (run under gpurun; one ncu capture per call)
set -e
export TFB200_FB_VARIANT=$1 PROF_ROWS=${4:-} PROF_REPS=2
python tools/prof_fb.py > gpurun_out/plain_$3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$2 -s ${NCU_SKIP:-9} -c 1 -f -o gpurun_out/$3 python tools/prof_fb.py > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
