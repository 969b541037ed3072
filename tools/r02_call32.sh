set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "scheduled_traversal or degenerate or api_contract" 2>&1 | tail -4
python bench.py --steps 14 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'])"
