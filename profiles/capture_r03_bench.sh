#!/bin/bash
# bench lines of every BASELINE workload (default lanes), and the reference arm for c3
T=${1:-r04z}
for w in c3 c3j c2 c4_dri c4_nodri c5_400 c5_440; do
  extra="--no-cpu-baseline"; [ $w = c3 ] && extra=""
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 $extra > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_ref_c3.json 2> gpurun_out/${T}_ref_c3.err
