set -x
CMD="python bench.py --steps 1 --profile"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3100 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:signsplit_group -s 585 -c 2 -o gpurun_out/prof_signsplit_r1 -f $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:transform_rotate_quant -s 590 -c 2 -o gpurun_out/prof_rotate_r1 -f $CMD > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fake_quant_group -s 295 -c 2 -o gpurun_out/prof_group_r1 -f $CMD > gpurun_out/ncu4.log 2>&1
echo finished
