set -x
P="$PWD/path-tracing...but-on-the-lumi-cluster_b200"
python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" > gpurun_out/r02_v8.log 2>&1
for v in cvt1 cvt2; do echo "== $v" >> gpurun_out/r02_v8.log; PTGPU_LIB=$P/libptgpu_$v.so python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1;flat=0" --check >> gpurun_out/r02_v8.log 2>&1; done
grep -v "^validate:" gpurun_out/r02_v8.log
