set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _d6; do
  echo "=== libptgpu$v.so (d6: bounce-ray sort key with 6 direction bits (octant, dominant axis, larger of the other two) and 7 origin bits instead of 3 + 10)" | tee -a gpurun_out/r02_ab41.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab41.log
done
