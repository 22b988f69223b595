#!/bin/bash
# weak scaling as the driver runs it: one rank per GPU (torchrun), the default bench (Panda headline + UR10 / Talos sub-records)
N=${1:-2}; TAG=${2:-r2}; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
  2> gpurun_out/scale_${TAG}_$N.err | tail -1 > gpurun_out/scale_${TAG}_$N.json
python -c "
import json;l=json.loads(open('gpurun_out/scale_${TAG}_$N.json').read());print('panda',l['n_gpus'],l['value'],l['ms_per_step'],l['e2e']['value'])
for k,v in l['extra']['workloads'].items(): print(k, v['value'], v['e2e']['value'])"
