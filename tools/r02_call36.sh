set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
PTGPU_LIB="$P/libptgpu_stats.so" timeout 900 python tools/stats_frames.py 520 1400 --validate > gpurun_out/r02_census36.log 2>&1; tail -60 gpurun_out/r02_census36.log
