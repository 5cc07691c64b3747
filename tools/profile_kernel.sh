#!/usr/bin/env bash
# one full-set ncu capture of one kernel from tools/kbench.py:  profile_kernel.sh <kbench-name> <kernel-regex> <tag>
ONLY="$1"; RE="$2"; TAG="$3"
export KB_ONLY="$ONLY" KB_ITERS=3
python tools/kbench.py > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$RE -s 6 -c 1 -o gpurun_out/prof_$TAG -f python tools/kbench.py > gpurun_out/ncu_$TAG.log 2>&1
