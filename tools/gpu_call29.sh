#!/bin/bash
# low-bit GEMM tests + the bench legs that use it (short bench)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_gemm_codes.py -x -q -m gpu 2>&1 | tail -8
echo "== bench (short)"; timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-e2e --no-generation --no-reference-legs --no-other-configs --no-cpu > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err; echo "rc $?"; tail -3 gpurun_out/c29_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/c29_bench.json') if l.startswith('{')][-1]
for k in ('value','ms_per_step','search','lowbit_gemm'):
    print(k, json.dumps(d.get(k))[:3000])
PY
