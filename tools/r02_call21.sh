set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wf_trace_cw" -c 2 -o gpurun_out/r02_src21 -f python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu21.log 2>&1
tail -3 gpurun_out/ncu21.log; ls -la gpurun_out/*.ncu-rep
