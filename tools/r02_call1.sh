set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; free -g | head -2
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer.log 2>&1; tail -5 gpurun_out/r02_sanitizer.log
python tools/ab_frames.py --check > gpurun_out/r02_ab1.log 2>&1; tail -40 gpurun_out/r02_ab1.log
for cfgname in base new; do
  if [ $cfgname = base ]; then OPTS="--opt flat=0 --opt sort=0"; else OPTS="--opt flat=1 --opt sort=1"; fi
  ncu --set full --clock-control none --import-source on -k regex:"wf_trace_cw" -c 3 -o gpurun_out/r02_full_$cfgname -f python tools/prof_frame.py --frames 520 --reps 1 --spp 64 $OPTS > gpurun_out/ncu_f_$cfgname.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches520_new.csv python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu_l.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; tail -15 gpurun_out/r02_pytest1.log
ls -la gpurun_out
