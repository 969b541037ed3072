set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_n4.json 2> gpurun_out/r02_scale_n4.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_scale_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['per_rank']['device_ms_per_step'], d['animation_seconds_measured'])"
