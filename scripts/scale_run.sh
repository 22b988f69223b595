#!/bin/bash
N=${1:-2}; mkdir -p gpurun_out
for w in talos panda; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --no-cpu-baseline 2> gpurun_out/scale_${w}_$N.err | tail -1 > gpurun_out/scale_${w}_$N.json
python -c "
import json;l=json.loads(open('gpurun_out/scale_${w}_$N.json').read());print('$w',l['n_gpus'],l['value'],l['ms_per_step'],l['e2e']['value'])"
done
