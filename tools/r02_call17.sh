set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
echo "=== lean state, no union box (libptgpu_u0.so)" | tee -a gpurun_out/r02_ab17.log
PTGPU_LIB="$P/libptgpu_u0.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab17.log
echo "=== lean state + union box (libptgpu.so)" | tee -a gpurun_out/r02_ab17.log
timeout 900 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1;flat=1,node_burst=4;flat=1,node_burst=5;flat=1,min_active=4;flat=1,min_active=8;flat=0" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab17.log
