#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 100 python tools/gemm_bringup.py --bench --short > gpurun_out/c35_bringup.log 2>&1; echo "rc $?"; grep -E "^d30" gpurun_out/c35_bringup.log | cut -c1-330
