set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in _b9 _b10; do
  echo "=== variant libptgpu$v.so"
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab15.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wf_trace_cw" -c 2 -o gpurun_out/r02_src15 -f python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu15.log 2>&1
tail -3 gpurun_out/ncu15.log; ls -la gpurun_out/*.ncu-rep
