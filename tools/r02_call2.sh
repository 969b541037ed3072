set -x
python tools/ab_frames.py --frames 0 330 520 1400 --configs "flat=1,sort=1;flat=1,sort=1,dyn_first=1;flat=1,sort=1,dyn_first=0,top_smem=1;flat=1,sort=1,dyn_first=1,top_smem=1" --check > gpurun_out/r02_ab2.log 2>&1; tail -30 gpurun_out/r02_ab2.log
PTGPU_LIB=$PWD/path-tracing...but-on-the-lumi-cluster_b200/libptgpu_b8.so python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1,sort=1;flat=1,sort=1,dyn_first=1" > gpurun_out/r02_ab2_b8.log 2>&1; tail -8 gpurun_out/r02_ab2_b8.log
PTGPU_LIB=$PWD/path-tracing...but-on-the-lumi-cluster_b200/libptgpu_stats.so python tools/stats_frames.py > gpurun_out/r02_stats_flat.log 2>&1; tail -20 gpurun_out/r02_stats_flat.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1; tail -30 gpurun_out/r02_pytest2.log
