#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== ubench"; timeout 120 fpqvar_b200/variants/ubench > gpurun_out/c2_ubench.log 2>&1; echo "rc $?"; cat gpurun_out/c2_ubench.log
export KB_ITERS=12 KB_NBUF=4
echo "== plain"; KB_ONLY="f16 (fc2)" timeout 200 python tools/kbench.py > gpurun_out/c2_plain_split.log 2>&1 && KB_ONLY="rotate+quant" timeout 200 python tools/kbench.py > gpurun_out/c2_plain_rot.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tma_kernel|signsplit_group_h16" -s 6 -c 2 -o gpurun_out/c2_rot KB_ONLY="rotate+quant" python tools/kbench.py > gpurun_out/c2_ncu_rot.log 2>&1
echo "rc $?"; tail -3 gpurun_out/c2_ncu_rot.log
