set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in _b9 _b9s7 _b9s8p3 _b9s12; do
  echo "=== 9 blocks/SM, 56 registers; libptgpu$v.so (sN: N stack entries in shared memory, pN: pending list depth; b9 = s10 p4)" | tee -a gpurun_out/r02_ab19.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab19.log
done
