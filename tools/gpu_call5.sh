#!/bin/bash
# ncu --set full of the two streaming rotate kernels (kbench shapes: 102400 x 1920) + the sign-split group kernel for reference
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export KB_ITERS=3 KB_NBUF=4
KB_ONLY="rotate+quant|adaLN+rotate" python tools/kbench.py > gpurun_out/c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stream_kernel" -s 8 -c 2 -o gpurun_out/c5_rot -f env KB_ONLY="rotate+quant|adaLN+rotate" python tools/kbench.py > gpurun_out/c5_ncu.log 2>&1
echo "rc $?"; cat gpurun_out/c5_plain.log; tail -5 gpurun_out/c5_ncu.log
