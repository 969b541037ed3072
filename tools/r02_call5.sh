set -x
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest5.log 2>&1; tail -25 gpurun_out/r02_pytest5.log
python tests/test_animation_gpu.py > gpurun_out/r02_animation_validation.md 2>gpurun_out/r02_animation_validation.err; cat gpurun_out/r02_animation_validation.md
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 4000 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1;flat=1,top_smem=1;flat=1,sort=0;flat=0;flat=0,sort=0" > gpurun_out/r02_ab5.log 2>&1; cat gpurun_out/r02_ab5.log
