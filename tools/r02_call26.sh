set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _tb2 _tb4 _fc128 _p5s6 _sky2; do
  echo "=== libptgpu$v.so (tb2/tb4: 2 or 4 triangle steps per TRI block (3); fc128: queue fetch chunk 128 (64); p5s6: pending list 5 deep + 6 stack entries in shared memory (4 + 7); sky2: sky march unrolled by 2)" | tee -a gpurun_out/r02_ab26.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab26.log
done
