set -x
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest23.log 2>&1; tail -15 gpurun_out/r02_pytest23.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench23.json 2> gpurun_out/r02_bench23.err; tail -c 3000 gpurun_out/r02_bench23.json; tail -3 gpurun_out/r02_bench23.err
