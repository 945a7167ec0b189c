#!/bin/bash
# ncu --set full of the fused IDCT + output kernels (c3, one lane)
export ROCJPEG_B200_LANES=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k23_' -s 6 -c 2 -o gpurun_out/${1:-r03e}_k23 $CMD > gpurun_out/${1:-r03e}_ncu.log 2>&1
