#!/bin/bash
# The reference's own jpegDecodePerf sample (built unmodified from /root/reference/samples, samples/Makefile)
# against this library, on the c3 image set written to a scratch directory. Run under gpurun.
set -e
D=$(mktemp -d)
python - "$D" <<'PY'
import sys
sys.path.insert(0, ".")
from rocjpeg_b200 import datagen
datas, _ = datagen.workload("c3", 256)
for i, d in enumerate(datas):
    open(f"{sys.argv[1]}/img_{i:04d}.jpg", "wb").write(d)
PY
for T in 1 4; do for B in 32 256; do
  echo "== jpegdecodeperf -fmt rgb_planar -t $T -b $B"
  samples/_build/jpegdecodeperf -i "$D" -fmt rgb_planar -t $T -b $B 2>&1 | grep -iE "images|time|per sec|fps|average|total" | head -12
done; done
rm -rf "$D"
