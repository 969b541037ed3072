set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in _tb4 _tb6 _tb16; do
  echo "=== libptgpu$v.so (tbN: up to N triangle steps per TRI block)" | tee -a gpurun_out/r02_ab27.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1;flat=1,tri_threshold=10" 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab27.log
done
