#!/bin/bash
# ncu --set full of the final low-bit GEMM: groups of 128, row scales, row scales with CTA pairs
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python tools/gemm_profile.py > gpurun_out/c37_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_codes -s 3 -c 3 -o gpurun_out/c37_gemm -f python tools/gemm_profile.py > gpurun_out/c37_ncu.log 2>&1
echo "rc $?"; tail -2 gpurun_out/c37_ncu.log
