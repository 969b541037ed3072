set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _m1 _m3 _m7; do
  echo "=== 9 blocks/SM kernel; byte->float conversions on the ALU pipe (PRMT, 2^15 bias): libptgpu$v.so (m1: x axis, m3: x and y, m7: all; none: I2F.U8 on the XU pipe)" | tee -a gpurun_out/r02_ab22.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab22.log
done
