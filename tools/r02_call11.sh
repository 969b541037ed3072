set -x
timeout 2400 python -m pytest tests -m gpu -q -rP > gpurun_out/r02_pytest11.log 2>&1; tail -5 gpurun_out/r02_pytest11.log; grep -i "within 1e-3 of\|golden frame 0:\|every 10th row\|GPU-vs-oracle\|validator frame\|frame .*: PSNR" gpurun_out/r02_pytest11.log | head -60
cd oracle/_ref && ./pt_gpu --gpus 1 --frames 0 40 --no-write 2>&1 | tail -8; cd ../..
