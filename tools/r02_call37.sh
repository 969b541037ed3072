set -x
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest37.log 2>&1; tail -4 gpurun_out/r02_pytest37.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
