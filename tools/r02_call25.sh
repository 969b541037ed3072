set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in _sp0 ""; do
  echo "=== shade kernels: one combined block append (3 barriers instead of 6); libptgpu$v.so (sp0: without, default: with the L2 prefetch of the next slot's path state)" | tee -a gpurun_out/r02_ab25.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab25.log
done
