#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/c15_gpu_tests.log 2>&1; echo "rc $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c15_gpu_tests.log | head -20 | cut -c1-300
echo "== kbench"; timeout 400 python tools/kbench.py 2>&1 | tee gpurun_out/c15_kbench.log
echo "== generation harness fp16 / fused"; timeout 600 python tools/var_generate.py --depth 30 --batch 50 --mode fp16,fused --iters 5 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['mode'], d['ms_per_batch'], d['value'], d.get('fpq_launches_per_batch'))"
echo "== reference model fused"; timeout 900 python tools/ref_model_generate.py --iters 3 --modes fp16,fused 2>&1 | grep "^{" | cut -c1-330
echo "== launch list of one profiled step"
python bench.py --profile --steps 1 --no-generation > gpurun_out/c15_profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/c15_launches.csv python bench.py --profile --steps 1 --no-generation > gpurun_out/c15_profile_ncu.log 2>&1
echo "rc $?"; tail -2 gpurun_out/c15_profile_plain.log | cut -c1-300
