set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in _na0 ""; do
  echo "=== 9 blocks/SM, 56 registers, 7 stack entries in shared memory; libptgpu$v.so (na0: triangle loads allocate in L1; default: ld.global.nc.L1::no_allocate)" | tee -a gpurun_out/r02_ab20.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab20.log
done
