#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/c22_gpu_tests.log 2>&1; echo "rc $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c22_gpu_tests.log | head -10 | cut -c1-300
echo "== kbench"; timeout 400 python tools/kbench.py 2>&1 | tee gpurun_out/c22_kbench.log
echo "== sustained"; for K in "signsplit f16 +clip" "group e2m1 f16" "rotate+quant" "adaLN+rotate"; do KB_ITERS=3000 KB_ONLY="$K" timeout 300 python tools/kbench.py 2>&1 | tail -2; done
echo "== stagebench"; for W in var_d30_w4a4_rot var_d30_w4a4_rot_nomod var_d36_w6a6_rot var_d16_w4a4; do WORKLOAD=$W timeout 300 python tools/stagebench.py > gpurun_out/c22_stage_$W.log 2>&1; tail -1 gpurun_out/c22_stage_$W.log; done
echo "== rowbench"; timeout 300 python tools/rowbench.py 2>&1 | tee gpurun_out/c22_rowbench.log | cut -c1-120
echo "== ncu dominant kernel + rotate"
export KB_ITERS=3 KB_NBUF=4
KB_ONLY="signsplit f16 +clip|adaLN+rotate" python tools/kbench.py > gpurun_out/c22_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"signsplit_group_h16|modulate_transform_rotate_quant_stream" -s 5 -c 3 -o gpurun_out/c22_final -f env KB_ONLY="signsplit f16 +clip|adaLN+rotate" python tools/kbench.py > gpurun_out/c22_ncu.log 2>&1
echo "rc $?"; tail -2 gpurun_out/c22_ncu.log
