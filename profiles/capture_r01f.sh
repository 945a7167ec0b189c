#!/bin/bash
# Round-1 (f, final) captures on c3 (default bench workload), one pipeline lane so that stage kernels do not
# overlap: launch list of the bench steps + ncu --set full of every stage kernel of one warmed-up step.
# Run under gpurun from the repo root; outputs land in gpurun_out/.
set -e
export ROCJPEG_B200_LANES=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r01f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches.csv $CMD > gpurun_out/r01f_ncu1.log 2>&1
$CMD > gpurun_out/r01f_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k1_sync|k1_write|k1_scan|k2_idct|k3_output|dc_' -s 20 -c 10 -o gpurun_out/r01f_prof $CMD > gpurun_out/r01f_ncu2.log 2>&1
