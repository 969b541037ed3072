set -x
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest29.log 2>&1; tail -4 gpurun_out/r02_pytest29.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -c 2500 gpurun_out/r02_bench_n1_final.json; tail -2 gpurun_out/r02_bench_n1_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2> gpurun_out/r02_bench_ref_final.err; tail -c 1200 gpurun_out/r02_bench_ref_final.json
python tools/prof_frame.py --frames 520 --reps 1 --spp 256 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches520_final.csv python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wf_trace_cw|wf_shade|wf_sort_scatter" -c 12 -o gpurun_out/r02_final2_full -f python tools/prof_frame.py --frames 520 --reps 1 --spp 256 > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/*.ncu-rep
