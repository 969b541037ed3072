set -x
echo "=== wf_trace_cw writes the shading-record index instead of the primitive (wf_shade<NEAR>'s gather does not wait for the instance record); traversal kernel: 22 B of spills" | tee -a gpurun_out/r02_ab48.log
timeout 900 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab48.log
