#!/bin/bash
# Round-1 (c) captures: launch list of one bench step + ncu --set full of every stage kernel (c3, one lane).
# Run under gpurun from the repo root; outputs land in gpurun_out/.
set -e
export ROCJPEG_B200_LANES=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r01c_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/r01c_ncu1.log 2>&1
$CMD > gpurun_out/r01c_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k1_sync|k1_write|k2_idct|k3_output|dc_' -s 14 -c 7 -o gpurun_out/r01c_prof $CMD > gpurun_out/r01c_ncu2.log 2>&1
