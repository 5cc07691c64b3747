#!/bin/bash
# ncu --set full of the low-bit GEMM (three variants)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python tools/gemm_profile.py > gpurun_out/c26_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_codes -s 3 -c 3 -o gpurun_out/c26_gemm -f python tools/gemm_profile.py > gpurun_out/c26_ncu.log 2>&1
echo "rc $?"; tail -5 gpurun_out/c26_ncu.log
