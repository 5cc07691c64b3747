#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu > gpurun_out/c13_gpu_tests.log 2>&1; echo "rc $?"; tail -12 gpurun_out/c13_gpu_tests.log | cut -c1-300
echo "== kbench gelu / score"; KB_ONLY="gelu|score|signsplit f16 +clip" timeout 300 python tools/kbench.py 2>&1 | tee gpurun_out/c13_kbench.log
