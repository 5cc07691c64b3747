#!/bin/bash
# Sweep of the shared-memory carveout every activation kernel asks for (tunable smem_kb).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export KB_ONLY="signsplit f16 (fc2)|group e2m1 f16|rotate+quant|adaLN+rotate|KV rows"
for KB in 228 196 164 132 100 64 0; do
  echo "== smem_kb=$KB"
  FPQ_TUNABLES=smem_kb=$KB timeout 200 python tools/kbench.py > gpurun_out/c4_kbench_$KB.log 2>&1; echo "rc $?"; cat gpurun_out/c4_kbench_$KB.log
  FPQ_TUNABLES=smem_kb=$KB WORKLOAD=var_d30_w4a4_rot timeout 200 python tools/stagebench.py > gpurun_out/c4_stage_$KB.log 2>&1; tail -1 gpurun_out/c4_stage_$KB.log
done
