#!/bin/bash
# 2 GPUs: the exhaustive self-tests of the new codes, the pinned-copy probe on both ranks, bench.py under torchrun
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== selftests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rotate_score.py -q -m gpu --timeout 120 2>&1 | tail -3
echo "== pcie probe 2 ranks (unbound / bound)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/pcie_probe.py 2>&1 | grep "^{"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/pcie_probe.py --bind 2>&1 | grep "^{"
nvidia-smi topo -m 2>&1 | head -12
echo "== bench 2 GPUs"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c16_bench2.json 2> gpurun_out/c16_bench2.err; echo "rc $?"; tail -3 gpurun_out/c16_bench2.err | cut -c1-300; python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/c16_bench2.json') if l.startswith('{')][-1]
for k in ('value','ms_per_step','n_gpus','e2e','search','generation','clocks'):
    print(k, json.dumps(d.get(k))[:700])
PY
