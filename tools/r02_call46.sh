set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _s64 _s256; do
  echo "=== libptgpu$v.so (s64 / s256: shade kernels in blocks of 64 / 256 threads instead of 128: the block-wide append barriers span 2 / 8 warps instead of 4)" | tee -a gpurun_out/r02_ab46.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab46.log
done
