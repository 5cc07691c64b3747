#!/usr/bin/env bash
# ncu evidence for profiles/: launch list of one step + full-set captures of the three hot kernels.
# Run under gpurun:  gpurun --timeout 1500 -- 'bash tools/profile_r1.sh <tag>'
TAG=${1:-r1}
CMD="python bench.py --steps 1 --profile"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1230 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
# the last-stage (largest) launch of each kernel in the warm-up replay: 30 blocks x {qkv, proj, fc1, fc2}
ncu --set full --clock-control none --import-source on -k regex:signsplit_group_h16 -s 292 -c 1 -o gpurun_out/prof_signsplit_$TAG -f $CMD > gpurun_out/ncu_s_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:transform_rotate_quant -s 590 -c 1 -o gpurun_out/prof_rotate_$TAG -f $CMD > gpurun_out/ncu_r_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fake_quant_group_h16 -s 295 -c 1 -o gpurun_out/prof_group_$TAG -f $CMD > gpurun_out/ncu_g_$TAG.log 2>&1
echo finished
