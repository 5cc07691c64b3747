#!/bin/bash
# cta_group::2 bring-up: the pair ladder alone first (short timeout), then the full bring-up + tests only if it passes
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== pair ladder"; timeout 45 python - <<'PY' > gpurun_out/c34_pair.log 2>&1
import sys, torch
sys.path.insert(0, ".")
from tools import gemm_bringup as B
from fpqvar_b200 import _lib as L
dev = torch.device("cuda:0")
L.set_tunable("gemm_pair", 1)
ok = B.check_gemm(dev)
print("PAIR LADDER", "OK" if ok else "FAILED")
PY
rc=$?; echo "rc $rc"; tail -12 gpurun_out/c34_pair.log
if [ $rc -eq 0 ] && grep -q "PAIR LADDER OK" gpurun_out/c34_pair.log; then bash tools/gpu_call24.sh; fi
