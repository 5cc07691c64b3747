#!/bin/bash
# full GPU suite, smoke, pack-kernel timing, then ncu --set full of the final low-bit GEMM
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full gpu suite"; T0=$(date +%s); timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6; echo "wall $(( $(date +%s) - T0 )) s"
echo "== bringup (short)"; timeout 300 python tools/gemm_bringup.py --bench --short > gpurun_out/c30_bringup.log 2>&1; echo "rc $?"; grep -E "^d30|pack parity" gpurun_out/c30_bringup.log | cut -c1-900
echo "== ncu"; python tools/gemm_profile.py > gpurun_out/c30_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_codes -s 3 -c 3 -o gpurun_out/c30_gemm -f python tools/gemm_profile.py > gpurun_out/c30_ncu.log 2>&1
echo "rc $?"; tail -2 gpurun_out/c30_ncu.log
