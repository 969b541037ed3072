set -x
P="$PWD/path-tracing...but-on-the-lumi-cluster_b200"
python tools/ab_frames.py --frames 0 330 520 1400 --configs "flat=1;flat=1,dyn_first=0" --check > gpurun_out/r02_ab3.log 2>&1; grep -v "^validate:" gpurun_out/r02_ab3.log | tail -12
python tools/ab_frames.py --frames 520 1400 --configs "flat=1;flat=1,tri_threshold=6;flat=1,tri_threshold=12;flat=1,node_threshold=12;flat=1,node_threshold=20;flat=1,min_active=4;flat=1,min_active=8;flat=1,min_active=12;flat=1,node_burst=3;flat=1,node_burst=4;flat=1,xform_threshold=2;flat=1,xform_threshold=8;flat=1,sort=0" > gpurun_out/r02_sweep3.log 2>&1; cat gpurun_out/r02_sweep3.log
for v in sm12 sm18 pend6 fetch32 b7; do echo "== $v" >> gpurun_out/r02_variants3.log; PTGPU_LIB=$P/libptgpu_$v.so python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" >> gpurun_out/r02_variants3.log 2>&1; done; cat gpurun_out/r02_variants3.log
python tests/test_subfunctions_gpu.py > gpurun_out/r02_subfn_measured.log 2>&1; cat gpurun_out/r02_subfn_measured.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest3.log 2>&1; tail -30 gpurun_out/r02_pytest3.log
