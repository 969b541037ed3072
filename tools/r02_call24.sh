set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
echo "=== libptgpu.so (9 blocks/SM, x conversions on the ALU pipe): scheduling thresholds" | tee -a gpurun_out/r02_ab24.log
timeout 600 python tools/ab_frames.py --frames 520 1400 --configs "flat=1;flat=1,node_threshold=12;flat=1,node_threshold=20;flat=1,tri_threshold=6;flat=1,tri_threshold=12;flat=1,min_active=8" 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab24.log
for v in _frcp _noadv _el; do
  echo "=== libptgpu$v.so (frcp: __frcp_rn instead of 1/det in the triangle test; noadv: no pop step before the ENTER vote; el: node loads with L1::evict_last)" | tee -a gpurun_out/r02_ab24.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab24.log
done
