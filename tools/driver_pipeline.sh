#!/bin/bash
# Serial vs pipelined host driver on one GPU: every 8th frame of the animation (225 frames), BMPs written.
set -e
cd oracle/_ref
mkdir -p /tmp/out_serial /tmp/out_pipe
./pt_gpu --gpus 1 --step 8 --serial --out /tmp/out_serial | tail -1
./pt_gpu --gpus 1 --step 8 --out /tmp/out_pipe | tail -1
n=0; bad=0
for f in /tmp/out_serial/*.bmp; do n=$((n+1)); cmp -s "$f" /tmp/out_pipe/$(basename "$f") || bad=$((bad+1)); done
echo "compared $n frames, $bad differ"
