#!/bin/bash
# Round-1 (h) ncu --set full of every stage kernel of one warmed-up step of the default bench command (c3, one lane).
set -e
export ROCJPEG_B200_LANES=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r01h_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k1_sync|k1_write|k2_idct|k3_output|dc_' -s 12 -c 6 -o gpurun_out/r01h_prof $CMD > gpurun_out/r01h_ncu2.log 2>&1
