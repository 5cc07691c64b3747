#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c6_gpu_tests.log 2>&1; echo "rc $?"; tail -5 gpurun_out/c6_gpu_tests.log
echo "== kbench"; KB_ONLY="rotate|signsplit f16 (fc2)" timeout 300 python tools/kbench.py 2>&1 | tee gpurun_out/c6_kbench.log
echo "== stagebench"; WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c6_stage_mod.log 2>&1; tail -1 gpurun_out/c6_stage_mod.log
WORKLOAD=var_d30_w4a4_rot_nomod timeout 300 python tools/stagebench.py > gpurun_out/c6_stage_nomod.log 2>&1; tail -1 gpurun_out/c6_stage_nomod.log
echo "== reference model"; timeout 900 python tools/ref_model_generate.py --iters 1 > gpurun_out/c6_refmodel.log 2>&1; echo "rc $?"; grep -v Warning gpurun_out/c6_refmodel.log | tail -12 | cut -c1-400
echo "== ncu"
export KB_ITERS=3 KB_NBUF=4
KB_ONLY="rotate+quant|adaLN+rotate" python tools/kbench.py > gpurun_out/c6_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stream_kernel" -s 6 -c 2 -o gpurun_out/c6_rot -f env KB_ONLY="rotate+quant|adaLN+rotate" python tools/kbench.py > gpurun_out/c6_ncu.log 2>&1
echo "rc $?"; tail -3 gpurun_out/c6_ncu.log
