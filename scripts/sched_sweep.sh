#!/bin/bash
# development aid: sweep of the launch-schedule knobs (loik_set_schedule, through scripts/lane_perf.py's SCHED variable)
# on one solve alone and on pipelined solves
R=${1:-panda}
for cfg in "dense_sweeps=4" "dense_sweeps=3" "dense_sweeps=5" "hi_priority_after=-1" "hi_priority_after=4" "repack_reps=3" "repack_reps=1" \
           "repack_growth=1.5" "drop_workspace=0" "lane_warps_per_cta=2"; do
  echo -n "$cfg: "; SCHED=$cfg LANE_AFTERS=${LANE_AFTERS:-32} DEPTHS=${DEPTHS:-1,10,32} python scripts/lane_perf.py $R 2>&1 | tail -1 | cut -c60-
done
