#!/bin/bash
# final state: smoke + full GPU suite
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full gpu suite"; T0=$(date +%s); timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4; echo "wall $(( $(date +%s) - T0 )) s"
