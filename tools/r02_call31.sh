set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err
tail -c 600 gpurun_out/r02_scale_n2.json | head -c 300; tail -2 gpurun_out/r02_scale_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_scale_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['per_rank'])"
