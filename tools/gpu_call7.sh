#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c7_gpu_tests.log 2>&1; echo "rc $?"; tail -5 gpurun_out/c7_gpu_tests.log
echo "== stagebench"; WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c7_stage_mod.log 2>&1; tail -1 gpurun_out/c7_stage_mod.log
echo "== reference model"; timeout 900 python tools/ref_model_generate.py --iters 2 > gpurun_out/c7_refmodel.log 2>&1; echo "rc $?"; grep "^{" gpurun_out/c7_refmodel.log | cut -c1-400
echo "== kv"; for M in "" "--quant-kv"; do timeout 600 python tools/var_generate.py --depth 30 --batch 50 --mode fused --iters 3 $M 2>&1 | grep "^{" | cut -c1-300; done
echo "== bench"; timeout 1500 python bench.py > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; echo "rc $?"; tail -3 gpurun_out/c7_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c7_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline','e2e','cpu_baseline','cpu_port','reference_gpu_path','generation','generation_reference_model','other_configs','clocks'):
    print(k, json.dumps(d.get(k))[:700])
PY
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c7_bench_ref.json 2> gpurun_out/c7_bench_ref.err; echo "rc $?"; cut -c1-900 gpurun_out/c7_bench_ref.json
