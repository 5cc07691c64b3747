#!/bin/bash
# A/B of the packed fp16 quantizer's rounding: magic-number FFMA (default build) vs FP4/FP6 conversion hardware (variant hw)
# vs the round-1 build; parity of the default build first.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
V=$PWD/fpqvar_b200/variants
echo "== ubench"; timeout 120 $V/ubench > gpurun_out/c3_ubench.log 2>&1; echo "rc $?"; cat gpurun_out/c3_ubench.log
echo "== full gpu suite (default build)"; timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c3_gpu_tests.log 2>&1; echo "rc $?"; tail -5 gpurun_out/c3_gpu_tests.log
echo "== kbench default | hw | r1"
timeout 300 python tools/kbench.py > gpurun_out/c3_kbench_magic.log 2>&1; echo "rc $?"
FPQ_LIB_PATH=$V/libfpq_b200_hw.so timeout 300 python tools/kbench.py > gpurun_out/c3_kbench_hw.log 2>&1; echo "rc $?"
FPQ_LIB_PATH=$V/libfpq_b200_r1.so timeout 300 python tools/kbench.py > gpurun_out/c3_kbench_r1.log 2>&1; echo "rc $?"
paste -d'|' gpurun_out/c3_kbench_magic.log gpurun_out/c3_kbench_hw.log gpurun_out/c3_kbench_r1.log | awk -F'|' '{printf "%-62s|%s|%s\n", substr($1,1,62), substr($2,33,30), substr($3,33,30)}'
echo "== rowbench default | hw"
timeout 300 python tools/rowbench.py > gpurun_out/c3_rowbench_magic.log 2>&1; echo "rc $?"
FPQ_LIB_PATH=$V/libfpq_b200_hw.so timeout 300 python tools/rowbench.py > gpurun_out/c3_rowbench_hw.log 2>&1; echo "rc $?"
paste -d'|' gpurun_out/c3_rowbench_magic.log gpurun_out/c3_rowbench_hw.log | cut -c1-200
for W in var_d30_w4a4_rot var_d30_w4a4_rot_nomod; do
  echo "== stagebench default $W"; WORKLOAD=$W timeout 300 python tools/stagebench.py > gpurun_out/c3_stage_magic_$W.log 2>&1; echo "rc $?"; tail -1 gpurun_out/c3_stage_magic_$W.log
done
echo "== stagebench r1 nomod"; FPQ_LIB_PATH=$V/libfpq_b200_r1.so WORKLOAD=var_d30_w4a4_rot_nomod timeout 300 python tools/stagebench.py > gpurun_out/c3_stage_r1_nomod.log 2>&1; tail -1 gpurun_out/c3_stage_r1_nomod.log
echo "== stagebench r1 mod"; FPQ_LIB_PATH=$V/libfpq_b200_r1.so WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c3_stage_r1_mod.log 2>&1; tail -1 gpurun_out/c3_stage_r1_mod.log
