set -x
echo "=== camera rays carry the initial path state implicitly: wf_generate does not write, wf_shade does not read attenuation / contribution / pending NEE at bounce 0" | tee -a gpurun_out/r02_ab43.log
timeout 900 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab43.log
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
