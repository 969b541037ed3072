set -x
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest44.log 2>&1; tail -3 gpurun_out/r02_pytest44.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final3.json 2> gpurun_out/r02_bench_n1_final3.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n1_final3.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_frame']['frac'], d['ms_per_step'], d['per_rank']['trace_ms_per_step'], d['cpu_baseline']['value'], d['animation_seconds_measured'])"
