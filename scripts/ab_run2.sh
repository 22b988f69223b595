#!/bin/bash
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest"; python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pool in 0 1; do export LOIK_WS_POOL=$pool
echo "== POOL=$pool"; PIPE=1 DEPTHS=32 python scripts/quick_perf.py panda,ur10 2>&1 | grep -v "^ *$" | tail -6
for r in panda ur10; do ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_iterate -s 10 -c 1 python scripts/quick_perf.py $r 2>&1 | grep -E "dram__|time_duration|hit_rate" | tr -s ' ' | tr '\n' ';'; echo; done
done
