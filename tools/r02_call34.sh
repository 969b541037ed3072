set -x
P="/root/repo/path-tracing...but-on-the-lumi-cluster_b200"
for v in "" _sp2; do
  echo "=== libptgpu$v.so (sp2: wf_shade<NEAR> also prefetches the NEXT slot's shading record: hit -> instance index offset -> 144-byte record)" | tee -a gpurun_out/r02_ab34.log
  PTGPU_LIB="$P/libptgpu$v.so" timeout 600 python tools/ab_frames.py --frames 0 520 1400 --configs "flat=1" --check 2>&1 | grep -v "^flat scene" | tee -a gpurun_out/r02_ab34.log
done
