#!/bin/bash
# Round-1 (h, final) bench lines: every BASELINE workload with the CPU baseline, and the reference arm on c3.
# Run under gpurun from the repo root; outputs land in gpurun_out/.
for w in c3 c3j c2 c4_dri c4_nodri c5_400 c5_440; do
  python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r01h_bench_$w.json 2> gpurun_out/r01h_bench_$w.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01h_ref_c3.json 2> gpurun_out/r01h_ref_c3.err
tail -c 600 gpurun_out/r01h_bench_c3.json
