set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _t9 _t10; do
  echo "=== libptgpu$v.so (t9: triangle-test constants S, axis in shared memory, 9 blocks/SM; t10: same with 10 blocks/SM, 48 registers, 82 B of spills, 5 stack entries + pending list 3 deep in shared memory)" | tee -a gpurun_out/r02_ab33.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab33.log
done
