set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _ea _t64 _t96; do
  echo "=== libptgpu$v.so (ea: scale x 1/d by an integer add on the exponent field, -3 instructions per node; t64 / t96: blocks of 64 / 96 threads, 18 / 12 per SM, instead of 128 x 9)" | tee -a gpurun_out/r02_ab39.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab39.log
done
