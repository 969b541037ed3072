set -x
cd oracle/_ref && mkdir -p /tmp/a8 && ./pt_gpu --gpus 8 --frames 0 160 --out /tmp/a8 2>&1 | tail -14; cd ../..
