set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _h2; do
  echo "=== libptgpu$v.so (h2: hit mask in slot order from predicates + one XOR-permutation, leaf mask in the node instead of meta bytes; 56 registers, 54 B of spills)" | tee -a gpurun_out/r02_ab35.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab35.log
done
