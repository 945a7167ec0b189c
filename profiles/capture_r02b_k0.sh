#!/bin/bash
# Round-2 (b): launch list of the default bench command (c3, one lane) and ncu --set full of the K0 kernels on c4_nodri (8 pictures).
set -e
export ROCJPEG_B200_LANES=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r02b_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_c3_launches.csv $CMD > gpurun_out/r02b_ncu1.log 2>&1
CMD4="python bench.py --workload c4_nodri --batch 8 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD4 > gpurun_out/r02b_plain4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k0_' -s 9 -c 3 -o gpurun_out/r02b_k0 $CMD4 > gpurun_out/r02b_ncu2.log 2>&1
