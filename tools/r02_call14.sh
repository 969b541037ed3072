set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _m1 _m3 _m7 _h _hm3; do
  echo "=== variant libptgpu$v.so"
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab14.log
done
