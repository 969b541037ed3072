set -x
python bench.py --steps 14 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
tail -c 3000 gpurun_out/bench_final.json
python tools/prof_frame.py --frames 520 --reps 1 --spp 256 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches520.csv python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wf_trace_cw|wf_shade|wf_classify|wf_generate" -c 8 -o gpurun_out/final_full -f python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out
