#!/bin/bash
# Round-2 (e): launch list of c3 (one lane) and ncu --set full of the K0 kernels on c4_dri (8 pictures).
export ROCJPEG_B200_LANES=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r02e_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_c3_launches.csv $CMD > gpurun_out/r02e_ncu1.log 2>&1
CMD4="python bench.py --workload c4_dri --batch 8 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD4 > gpurun_out/r02e_plain4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k0_' -s 9 -c 3 -o gpurun_out/r02e_k0 $CMD4 > gpurun_out/r02e_ncu2.log 2>&1
