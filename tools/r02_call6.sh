set -x
# final kernels: launch list + full ncu capture of the traversal kernel and the shade kernels (frame 520)
python tools/prof_frame.py --frames 520 --reps 1 --spp 256 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches520.csv python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wf_trace_cw|wf_shade|wf_sort_scatter" -c 12 -o gpurun_out/r02_final_full -f python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/*.ncu-rep
# the whole animation on one GPU through the drop-in driver, frames written and scanned
cd oracle/_ref && mkdir -p /tmp/anim1 && ./pt_gpu --gpus 1 --out /tmp/anim1 > ../../gpurun_out/r02_full_animation_1gpu.log 2>&1; cd ../..
tail -5 gpurun_out/r02_full_animation_1gpu.log
python tools/scan_frames.py /tmp/anim1 > gpurun_out/r02_full_animation_1gpu_scan.log 2>&1; cat gpurun_out/r02_full_animation_1gpu_scan.log
