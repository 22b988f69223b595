#!/bin/bash
# development aid: sweep of the launch-schedule knobs (env) on the pipelined solve
export CUDA_DEVICE_MAX_CONNECTIONS=32
R=${1:-panda}
for cfg in "LOIK_DENSE=4" "LOIK_DENSE=3" "LOIK_DENSE=5" "LOIK_HI_AFTER=-1" "LOIK_HI_AFTER=4" "LOIK_HI_AFTER=16" "LOIK_REPS=3" "LOIK_REPS=1" "LOIK_GROWTH=1.5" "LOIK_REPS=3 LOIK_GROWTH=1.5" "LOIK_DENSE=5 LOIK_REPS=3"; do
  echo -n "$cfg: "; env $cfg PIPE=1 DEPTHS=32 python scripts/quick_perf.py $R 2>&1 | grep pipeline
done
