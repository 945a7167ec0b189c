#!/bin/bash
# ncu --set full of the IDCT kernel on the 4K workload (one lane)
export ROCJPEG_B200_LANES=1
CMD="python bench.py --workload c4_dri --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_idct' -s 4 -c 1 -o gpurun_out/${1:-r03l}_k2 $CMD > gpurun_out/${1:-r03l}_ncu.log 2>&1
