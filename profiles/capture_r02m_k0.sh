#!/bin/bash
# Round-2 (m): ncu --set full of the K0 kernels on c4_nodri (8 pictures), after the branch-light rewrite.
export ROCJPEG_B200_LANES=1
CMD4="python bench.py --workload c4_nodri --batch 8 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD4 > gpurun_out/r02m_plain4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k0_' -s 9 -c 3 -o gpurun_out/r02m_k0 $CMD4 > gpurun_out/r02m_ncu2.log 2>&1
