#!/bin/bash
# round-2 final bench lines as the driver runs them: the B200 arm and the reference arm, 1 GPU; launch list of one profiled step
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== bench"; T0=$(date +%s); timeout 2400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c32_bench.json 2> gpurun_out/c32_bench.err; echo "rc $? wall $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/c32_bench.err; python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/c32_bench.json') if l.startswith('{')][-1]
for k in ('value','ms_per_step','roofline','e2e','cpu_baseline','generation','generation_reference_model','search','lowbit_gemm','clocks','gpu_launches'):
    print(k, json.dumps(d.get(k))[:1200])
PY
echo "== reference arm"; T0=$(date +%s); timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/c32_bench_ref.json 2> gpurun_out/c32_bench_ref.err; echo "rc $? wall $(( $(date +%s) - T0 )) s"; cut -c1-800 gpurun_out/c32_bench_ref.json
