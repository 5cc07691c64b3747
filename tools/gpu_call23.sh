#!/bin/bash
# round-2 bench lines as the driver runs them: the B200 arm and the reference arm, 1 GPU
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== launch list"; python bench.py --profile --steps 1 --no-generation > gpurun_out/c23_profile_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/c23_launches.csv python bench.py --profile --steps 1 --no-generation > gpurun_out/c23_profile_ncu.log 2>&1; echo "rc $?"
echo "== bench"; T0=$(date +%s); timeout 2400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err; echo "rc $? wall $(( $(date +%s) - T0 )) s"; python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/c23_bench.json') if l.startswith('{')][-1]
for k in ('value','ms_per_step','roofline','kernels','e2e','cpu_baseline','cpu_port','reference_gpu_path','generation','generation_reference_model','search','other_configs','config1','clocks','gpu_launches'):
    print(k, json.dumps(d.get(k))[:1500])
PY
echo "== reference arm"; T0=$(date +%s); timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/c23_bench_ref.json 2> gpurun_out/c23_bench_ref.err; echo "rc $? wall $(( $(date +%s) - T0 )) s"; cut -c1-1200 gpurun_out/c23_bench_ref.json
