#!/bin/bash
# One GPU session: parity first (short timeouts: a hung kernel must not eat the session), then A/B measurements of this
# build against the round-1 build (fpqvar_b200/variants/libfpq_b200_r1.so).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
R1=$PWD/fpqvar_b200/variants/libfpq_b200_r1.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.log 2>&1
echo "== rotate streaming parity"; timeout 300 python -m pytest tests/test_gpu_shapes.py -x -q -m gpu -k "kernel_choice or without_pdl" > gpurun_out/c1_rot_parity.log 2>&1; echo "rc $?"; tail -3 gpurun_out/c1_rot_parity.log
echo "== exhaustive f16 flow"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/c1_parity.log 2>&1; echo "rc $?"; tail -3 gpurun_out/c1_parity.log
echo "== full gpu suite"; timeout 900 python -m pytest tests -q -m gpu > gpurun_out/c1_gpu_tests.log 2>&1; echo "rc $?"; tail -15 gpurun_out/c1_gpu_tests.log
echo "== kbench new"; KB_ONLY= timeout 300 python tools/kbench.py > gpurun_out/c1_kbench_new.log 2>&1; echo "rc $?"
echo "== kbench r1"; FPQ_LIB_PATH=$R1 timeout 300 python tools/kbench.py > gpurun_out/c1_kbench_r1.log 2>&1; echo "rc $?"
paste -d'|' gpurun_out/c1_kbench_new.log gpurun_out/c1_kbench_r1.log | cut -c1-170
for W in var_d30_w4a4_rot var_d30_w4a4_rot_nomod; do
  echo "== stagebench new $W"; WORKLOAD=$W timeout 300 python tools/stagebench.py > gpurun_out/c1_stage_new_$W.log 2>&1; echo "rc $?"; tail -1 gpurun_out/c1_stage_new_$W.log
  echo "== stagebench new $W all-streaming"; FPQ_TUNABLES=rot_small_max_chunks=0 WORKLOAD=$W timeout 300 python tools/stagebench.py > gpurun_out/c1_stage_new_stream_$W.log 2>&1; echo "rc $?"; tail -1 gpurun_out/c1_stage_new_stream_$W.log
  echo "== stagebench r1 $W"; FPQ_LIB_PATH=$R1 WORKLOAD=$W timeout 300 python tools/stagebench.py > gpurun_out/c1_stage_r1_$W.log 2>&1; echo "rc $?"; tail -1 gpurun_out/c1_stage_r1_$W.log
done
echo "== bench new"; timeout 600 python bench.py --no-generation > gpurun_out/c1_bench_new.json 2> gpurun_out/c1_bench_new.err; echo "rc $?"; cut -c1-1500 gpurun_out/c1_bench_new.json
echo "== bench r1 (nomod step)"; FPQ_LIB_PATH=$R1 timeout 600 python bench.py --no-generation --no-e2e --no-cpu --workload var_d30_w4a4_rot_nomod > gpurun_out/c1_bench_r1_nomod.json 2> gpurun_out/c1_bench_r1.err; echo "rc $?"; cut -c1-600 gpurun_out/c1_bench_r1_nomod.json
echo "== bench new (nomod step)"; timeout 600 python bench.py --no-generation --no-e2e --no-cpu --workload var_d30_w4a4_rot_nomod > gpurun_out/c1_bench_new_nomod.json 2> gpurun_out/c1_bench_new_nomod.err; echo "rc $?"; cut -c1-600 gpurun_out/c1_bench_new_nomod.json
