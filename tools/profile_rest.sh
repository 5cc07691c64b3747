#!/usr/bin/env bash
# one full-set ncu pass over the kernels that are not in the default bench step (scoring, adaLN-fused rotate, weight
# rotate, fp32 group, per-token rows, KV rows of 64, quant_cuda.quant compat):  gpurun -- 'bash tools/profile_rest.sh'
# The report stays on the box (it exceeds the 64 MiB that come back); only the raw-page CSV is returned.
export KB_ITERS=1 KB_NBUF=2
python tools/kbench.py > gpurun_out/plain_rest.log 2>&1 &&
ncu --set full --clock-control none \
    -k regex:"score_formats|transform_rotate_quant_kernel|transform_rotate_weight|quant_grid_kernel|fake_quant_group_kernel|fake_quant_row_reg" \
    -c 30 -o /tmp/prof_rest -f python tools/kbench.py > gpurun_out/ncu_rest.log 2>&1
echo "rc $?" >> gpurun_out/ncu_rest.log
ncu -i /tmp/prof_rest.ncu-rep --page raw --csv > gpurun_out/rest_raw.csv 2>> gpurun_out/ncu_rest.log
ls -la /tmp/prof_rest.ncu-rep >> gpurun_out/ncu_rest.log
