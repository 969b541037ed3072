set -x
nvidia-smi -L | head -8; nproc; free -g | head -2
for N in 1 2 8; do
  if [ $N = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err; fi
  tail -c 1500 gpurun_out/r02_scale_n$N.json; tail -2 gpurun_out/r02_scale_n$N.err
done
cd oracle/_ref && mkdir -p /tmp/anim8 && ./pt_gpu --gpus 8 --out /tmp/anim8 > ../../gpurun_out/r02_full_animation_8gpu.log 2>&1; cd ../..
tail -5 gpurun_out/r02_full_animation_8gpu.log
python tools/scan_frames.py /tmp/anim8 > gpurun_out/r02_full_animation_8gpu_scan.log 2>&1; cat gpurun_out/r02_full_animation_8gpu_scan.log
