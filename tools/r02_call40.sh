set -x
echo "=== L2 access-policy window over the BVH nodes (persisting), percent of the persisting-L2 maximum set aside" | tee -a gpurun_out/r02_ab40.log
timeout 900 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1;flat=1,l2_persist=25;flat=1,l2_persist=50;flat=1,l2_persist=100;flat=1" 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab40.log
