set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _b9 _b10; do
  echo "=== hit record in stack storage; libptgpu$v.so (b9: 9 blocks/SM 56 regs smem stack 10; b10: 10 blocks 48 regs smem stack 8)" | tee -a gpurun_out/r02_ab18.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab18.log
done
