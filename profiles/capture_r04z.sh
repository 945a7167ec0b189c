#!/bin/bash
# Round-2 final captures: launch list of the default bench command (default lanes), ncu --set full of every stage kernel
# on c3 and c4_dri (one lane, so that one step's launches come in order).
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r04z_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r04z_c3_launches.csv $CMD > gpurun_out/r04z_ncu1.log 2>&1
export ROCJPEG_B200_LANES=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k0_|k1_|dc_|k2_|k23_|k3_' -s 22 -c 11 -o gpurun_out/r04z_c3 $CMD > gpurun_out/r04z_ncu2.log 2>&1
CMD4="python bench.py --workload c4_dri --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD4 > gpurun_out/r04z_plain4.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none -k regex:'k0_|k1_|dc_|k2_|k23_|k3_' -s 22 -c 11 -o gpurun_out/r04z_c4 $CMD4 > gpurun_out/r04z_ncu4.log 2>&1
