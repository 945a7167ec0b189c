#!/bin/bash
# Round-1 (h) ncu launch list of the default bench command (c3), one pipeline lane so that the stage kernels
# do not overlap. One ncu pass per gpurun call, after the same command ran clean without ncu.
set -e
export ROCJPEG_B200_LANES=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r01h_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01h_launches.csv $CMD > gpurun_out/r01h_ncu1.log 2>&1
