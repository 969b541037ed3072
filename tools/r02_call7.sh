set -x
P="$PWD/path-tracing...but-on-the-lumi-cluster_b200"
python tools/ab_frames.py --frames 520 1400 --configs "flat=1" > gpurun_out/r02_v7.log 2>&1
for v in key15; do echo "== $v" >> gpurun_out/r02_v7.log; PTGPU_LIB=$P/libptgpu_$v.so python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" >> gpurun_out/r02_v7.log 2>&1; done
for v in pend8sm10 pend6sm12; do echo "== $v" >> gpurun_out/r02_v7.log; PTGPU_LIB=$P/libptgpu_$v.so python tools/ab_frames.py --frames 520 1400 --configs "flat=1;flat=1,tri_threshold=12;flat=1,tri_threshold=16;flat=1,tri_threshold=20" >> gpurun_out/r02_v7.log 2>&1; done
cat gpurun_out/r02_v7.log
