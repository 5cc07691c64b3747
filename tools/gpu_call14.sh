#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== rotate_score tests"; timeout 600 python -m pytest tests/test_gpu_rotate_score.py tests/test_reference_search.py tests/test_gpu_dropin.py -q -m gpu --timeout 90 > gpurun_out/c14_tests.log 2>&1; echo "rc $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|assert " gpurun_out/c14_tests.log | head -30 | cut -c1-300
echo "== stagebench default / small-kernel threshold 60000"; WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c14_stage.log 2>&1; tail -1 gpurun_out/c14_stage.log
FPQ_TUNABLES=rot_small_max_chunks=60000 WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c14_stage_small60k.log 2>&1; tail -1 gpurun_out/c14_stage_small60k.log
paste -d'|' gpurun_out/c14_stage.log gpurun_out/c14_stage_small60k.log | grep -E "mat_qkv" | awk -F'|' '{printf "%s | %s\n", $1, substr($2,28,25)}'
