set -x
nvidia-smi -L | head -8; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
tail -c 1500 gpurun_out/r02_scale_n8.json; tail -2 gpurun_out/r02_scale_n8.err
cd oracle/_ref && mkdir -p /tmp/anim8 && ./pt_gpu --gpus 8 --out /tmp/anim8 > ../../gpurun_out/r02_full_animation_8gpu.log 2>&1; cd ../..
tail -5 gpurun_out/r02_full_animation_8gpu.log
python tools/scan_frames.py /tmp/anim8 > gpurun_out/r02_full_animation_8gpu_scan.log 2>&1; tail -5 gpurun_out/r02_full_animation_8gpu_scan.log
