#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
timeout 200 python -m pytest tests/test_gpu_gemm_codes.py -x -q -m gpu 2>&1 | tail -5
