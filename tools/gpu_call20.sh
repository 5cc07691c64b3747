#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
V=$PWD/fpqvar_b200/variants
echo "== gpu tests (default build)"; timeout 900 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/c20_gpu_tests.log 2>&1; echo "rc $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c20_gpu_tests.log | head -10 | cut -c1-300
echo "== signsplit tests on the splithw build"; FPQ_LIB_PATH=$V/libfpq_b200_splithw.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_ref_ext.py -q -m gpu --timeout 120 2>&1 | tail -3 | cut -c1-300
echo "== kbench default | splithw"
KB_ONLY="signsplit|rows int|gelu+" timeout 300 python tools/kbench.py > gpurun_out/c20_kb_def.log 2>&1
FPQ_LIB_PATH=$V/libfpq_b200_splithw.so KB_ONLY="signsplit|rows int|gelu+" timeout 300 python tools/kbench.py > gpurun_out/c20_kb_hw.log 2>&1
paste -d'|' gpurun_out/c20_kb_def.log gpurun_out/c20_kb_hw.log | cut -c1-150
echo "== sustained (KB_ITERS=3000)"
KB_ITERS=3000 KB_ONLY="signsplit f16 +clip" timeout 300 python tools/kbench.py 2>&1 | tail -2
FPQ_LIB_PATH=$V/libfpq_b200_splithw.so KB_ITERS=3000 KB_ONLY="signsplit f16 +clip" timeout 300 python tools/kbench.py 2>&1 | tail -2
for i in 1 2; do
echo "== stagebench default"; WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c20_stage_def_$i.log 2>&1; tail -1 gpurun_out/c20_stage_def_$i.log
echo "== stagebench splithw"; FPQ_LIB_PATH=$V/libfpq_b200_splithw.so WORKLOAD=var_d30_w4a4_rot timeout 300 python tools/stagebench.py > gpurun_out/c20_stage_hw_$i.log 2>&1; tail -1 gpurun_out/c20_stage_hw_$i.log
done
paste -d'|' gpurun_out/c20_stage_def_2.log gpurun_out/c20_stage_hw_2.log | grep fc2 | awk -F'|' '{printf "%s | %s\n", $1, substr($2,28,25)}'
echo "== bench value (no extras) default | splithw"
timeout 600 python bench.py --no-e2e --no-cpu --no-generation --no-reference-legs --no-other-configs --no-search 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k[:20]:round(v['GB/s']) for k,v in d['kernels'].items()}, d['clocks'])"
FPQ_LIB_PATH=$V/libfpq_b200_splithw.so timeout 600 python bench.py --no-e2e --no-cpu --no-generation --no-reference-legs --no-other-configs --no-search 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k[:20]:round(v['GB/s']) for k,v in d['kernels'].items()}, d['clocks'])"
