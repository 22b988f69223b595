#!/bin/bash
# development aid: A/B builds of the library on the GPU box (gpurun).  usage: ab_run.sh "<lib suffixes>" "<robots>"
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
LIBS=${1:-"base v1 v2 v3"}; ROBOTS=${2:-"panda,ur10,talos"}
for lib in $LIBS; do
  export LOIK_B200_LIB=$PWD/loik_b200/libloik_b200_$lib.so
  echo "== pytest $lib" ; python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  echo "== quick_perf $lib"; PIPE=1 DEPTHS=16 python scripts/quick_perf.py $ROBOTS 2>&1 | grep -v "^ *$" | tail -12
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio --clock-control none -k regex:k_iterate -s 10 -c 1 python scripts/quick_perf.py panda 2>&1 | grep -E "inst_executed|time_duration|stalled"
done
