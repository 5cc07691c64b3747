#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c11_gpu_tests.log 2>&1; echo "rc $?"; tail -5 gpurun_out/c11_gpu_tests.log
echo "== host overhead"; timeout 300 python tools/hostoverhead.py 2>&1 | tail -12
echo "== reference model"; timeout 1200 python tools/ref_model_generate.py --iters 2 > gpurun_out/c11_refmodel.log 2>&1; echo "rc $?"; grep "^{" gpurun_out/c11_refmodel.log | cut -c1-400
echo "== generation harness fp16 / fused / kv"; timeout 600 python tools/var_generate.py --depth 30 --batch 50 --mode fp16,fused --iters 3 2>&1 | grep "^{" | cut -c1-260
for M in incremental reference; do timeout 600 python tools/var_generate.py --depth 30 --batch 50 --mode fused --iters 3 --quant-kv $M 2>&1 | grep "^{" | cut -c1-260; done
echo "== bench"; timeout 2400 python bench.py > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; echo "rc $?"; tail -3 gpurun_out/c11_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c11_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline','e2e','cpu_baseline','cpu_port','reference_gpu_path','generation','generation_reference_model','other_configs','clocks'):
    print(k, json.dumps(d.get(k))[:900])
PY
