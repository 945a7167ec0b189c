#!/bin/bash
# ncu --set full of the stage kernels on the 4K 4:2:2 workload (8 pictures, one lane), after the round-1 K2 / K1 changes.
set -e
export ROCJPEG_B200_LANES=1
CMD="python bench.py --workload c4_nodri --batch 8 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r01g_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k1_sync|k1_write|k2_idct|dc_' -s 16 -c 8 -o gpurun_out/r01g_c4_prof $CMD > gpurun_out/r01g_ncu.log 2>&1
