#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c12_gpu_tests.log 2>&1; echo "rc $?"; tail -5 gpurun_out/c12_gpu_tests.log
echo "== kv (5 iters each, twice)"
for R in 1 2; do for M in "" "--quant-kv incremental" "--quant-kv reference"; do timeout 600 python tools/var_generate.py --depth 30 --batch 50 --mode fused --iters 5 $M 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$M', d['ms_per_batch'])"; done; done
echo "== bench (search leg)"; timeout 1200 python bench.py --no-reference-legs --no-other-configs --no-generation --no-cpu > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "rc $?"; tail -3 gpurun_out/c12_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c12_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','search','clocks'):
    print(k, json.dumps(d.get(k))[:1200])
PY
echo "== ncu dominant kernel"
export KB_ITERS=3 KB_NBUF=4
KB_ONLY="signsplit f16 +clip" python tools/kbench.py > gpurun_out/c12_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"signsplit_group_h16" -s 5 -c 1 -o gpurun_out/c12_split -f env KB_ONLY="signsplit f16 +clip" python tools/kbench.py > gpurun_out/c12_ncu.log 2>&1
echo "rc $?"; cat gpurun_out/c12_plain.log; tail -3 gpurun_out/c12_ncu.log
