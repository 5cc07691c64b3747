#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
V=$PWD/fpqvar_b200/variants
echo "== gpu tests (default build: sign-split on the conversion hardware)"; timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/c21_gpu_tests.log 2>&1; echo "rc $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c21_gpu_tests.log | head -10 | cut -c1-300
echo "== kbench default | symhw"
KB_ONLY="group e2m|KV rows|rows e2m3" timeout 300 python tools/kbench.py > gpurun_out/c21_kb_def.log 2>&1
FPQ_LIB_PATH=$V/libfpq_b200_symhw.so KB_ONLY="group e2m|KV rows|rows e2m3" timeout 300 python tools/kbench.py > gpurun_out/c21_kb_hw.log 2>&1
paste -d'|' gpurun_out/c21_kb_def.log gpurun_out/c21_kb_hw.log | cut -c1-150
echo "== sustained group e2m1 (KB_ITERS=3000)"
KB_ITERS=3000 KB_ONLY="group e2m1 f16" timeout 300 python tools/kbench.py 2>&1 | tail -2
FPQ_LIB_PATH=$V/libfpq_b200_symhw.so KB_ITERS=3000 KB_ONLY="group e2m1 f16" timeout 300 python tools/kbench.py 2>&1 | tail -2
for i in 1 2; do
echo "== stagebench default"; WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c21_stage_def_$i.log 2>&1; tail -1 gpurun_out/c21_stage_def_$i.log
echo "== stagebench symhw"; FPQ_LIB_PATH=$V/libfpq_b200_symhw.so WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c21_stage_hw_$i.log 2>&1; tail -1 gpurun_out/c21_stage_hw_$i.log
done
echo "== bench value (no extras) default | symhw"
timeout 600 python bench.py --no-e2e --no-cpu --no-generation --no-reference-legs --no-other-configs --no-search 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k[:20]:round(v['GB/s']) for k,v in d['kernels'].items()}, d['clocks'], d['roofline']['frac'])"
FPQ_LIB_PATH=$V/libfpq_b200_symhw.so timeout 600 python bench.py --no-e2e --no-cpu --no-generation --no-reference-legs --no-other-configs --no-search 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k[:20]:round(v['GB/s']) for k,v in d['kernels'].items()}, d['clocks'], d['roofline']['frac'])"
echo "== d16 / d36 default | symhw"
for W in var_d16_w4a4 var_d36_w6a6_rot; do
WORKLOAD=$W timeout 300 python tools/stagebench.py 2>&1 | tail -1
FPQ_LIB_PATH=$V/libfpq_b200_symhw.so WORKLOAD=$W timeout 300 python tools/stagebench.py 2>&1 | tail -1
done
