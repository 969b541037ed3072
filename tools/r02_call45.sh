set -x
timeout 900 python tools/perf_configs.py 2>&1 | tee gpurun_out/r02_other_configs.log
