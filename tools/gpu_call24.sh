#!/bin/bash
# low-bit GEMM bring-up: parity ladder + timings, then the new GPU tests
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== bringup"; timeout 150 python tools/gemm_bringup.py --bench --short > gpurun_out/c24_bringup.log 2>&1; echo "rc $?"; tail -40 gpurun_out/c24_bringup.log
echo "== tests"; timeout 240 python -m pytest tests/test_gpu_gemm_codes.py -x -q -m gpu 2>&1 | tail -15
