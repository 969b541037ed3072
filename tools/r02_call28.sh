set -x
echo "=== plain single-ray loop (one thread per ray) against the warp-scheduled kernel on the same sorted queues: plain_trace 2 = primary round only, 1 = every round" | tee -a gpurun_out/r02_ab28.log
timeout 900 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1;flat=1,plain_trace=2;flat=1,plain_trace=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab28.log
