set -x
cd oracle/_ref && mkdir -p /tmp/anim8 && ./pt_gpu --gpus 8 --out /tmp/anim8 > ../../gpurun_out/r02_full_animation_8gpu.log 2>&1; cd ../..
tail -16 gpurun_out/r02_full_animation_8gpu.log
python tools/scan_frames.py /tmp/anim8 > gpurun_out/r02_full_animation_8gpu_scan.log 2>&1; cat gpurun_out/r02_full_animation_8gpu_scan.log
