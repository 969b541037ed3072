set -x
python tests/test_animation_gpu.py > gpurun_out/r02_animation_validation.md 2> gpurun_out/r02_animation_validation.err; tail -30 gpurun_out/r02_animation_validation.md; tail -3 gpurun_out/r02_animation_validation.err
