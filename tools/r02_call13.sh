set -x
cd oracle/_ref && mkdir -p /tmp/anim1 && ./pt_gpu --gpus 1 --out /tmp/anim1 > ../../gpurun_out/r02_full_animation_1gpu.log 2>&1; cd ../..
tail -8 gpurun_out/r02_full_animation_1gpu.log
python tools/scan_frames.py /tmp/anim1 > gpurun_out/r02_full_animation_1gpu_scan.log 2>&1; cat gpurun_out/r02_full_animation_1gpu_scan.log
