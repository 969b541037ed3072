#!/bin/bash
# Whole 1800-frame animation through the drop-in driver on all GPUs of the box, BMPs written to /tmp.
N=${1:-8}
cd oracle/_ref
mkdir -p /tmp/anim
./pt_gpu --gpus $N --out /tmp/anim | tail -3
ls /tmp/anim | wc -l
python3 - <<'PY'
import glob
fs = sorted(glob.glob('/tmp/anim/frame_*.bmp'))
vals = []
for f in fs[::100]:
    b = open(f, 'rb').read()
    px = b[54:]
    vals.append(round(sum(px[::97]) / max(1, len(px[::97]))))
print("mean brightness of every 100th frame:", vals)
PY
