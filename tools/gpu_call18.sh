#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
V=$PWD/fpqvar_b200/variants
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/c18_gpu_tests.log 2>&1; echo "rc $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c18_gpu_tests.log | head -20 | cut -c1-300
echo "== kbench rotate: default (hw quant + packed modulate) | rotmagic"
KB_ONLY="rotate" timeout 300 python tools/kbench.py > gpurun_out/c18_kb_hw.log 2>&1
FPQ_LIB_PATH=$V/libfpq_b200_rotmagic.so KB_ONLY="rotate" timeout 300 python tools/kbench.py > gpurun_out/c18_kb_magic.log 2>&1
paste -d'|' gpurun_out/c18_kb_hw.log gpurun_out/c18_kb_magic.log | cut -c1-140
for i in 1 2; do
echo "== stagebench default"; WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c18_stage_hw_$i.log 2>&1; tail -1 gpurun_out/c18_stage_hw_$i.log
echo "== stagebench rotmagic"; FPQ_LIB_PATH=$V/libfpq_b200_rotmagic.so WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c18_stage_magic_$i.log 2>&1; tail -1 gpurun_out/c18_stage_magic_$i.log
done
echo "== d36"; WORKLOAD=var_d36_w6a6_rot timeout 300 python tools/stagebench.py > gpurun_out/c18_stage_d36_hw.log 2>&1; tail -1 gpurun_out/c18_stage_d36_hw.log
FPQ_LIB_PATH=$V/libfpq_b200_rotmagic.so WORKLOAD=var_d36_w6a6_rot timeout 300 python tools/stagebench.py > gpurun_out/c18_stage_d36_magic.log 2>&1; tail -1 gpurun_out/c18_stage_d36_magic.log
