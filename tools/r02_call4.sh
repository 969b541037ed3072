set -x
P="$PWD/path-tracing...but-on-the-lumi-cluster_b200"
python tools/diag_dropin.py > gpurun_out/r02_diag_dropin.log 2>&1; cat gpurun_out/r02_diag_dropin.log
python tests/test_subfunctions_gpu.py > gpurun_out/r02_subfn_measured.log 2>&1; cat gpurun_out/r02_subfn_measured.log
PTGPU_LIB=$P/libptgpu_stats.so python tools/stats_frames.py --validate 520 1400 > gpurun_out/r02_stats_plain.log 2>&1; grep -v "^validate" gpurun_out/r02_stats_plain.log | tail -20
python tools/ab_frames.py --frames 520 1400 --configs "flat=1;flat=1,node_burst=3,xform_threshold=2;flat=1,node_burst=3,xform_threshold=1;flat=1,node_burst=3,xform_threshold=2,min_active=8" > gpurun_out/r02_sweep4.log 2>&1; cat gpurun_out/r02_sweep4.log
