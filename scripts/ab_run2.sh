#!/bin/bash
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
for lib in tmaA tmaB; do export LOIK_B200_LIB=$PWD/loik_b200/libloik_b200_$lib.so
echo "== pytest $lib"; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== perf $lib"; PIPE=1 DEPTHS=32 timeout 300 python scripts/quick_perf.py panda,ur10,talos 2>&1 | grep -v "^ *$" | tail -9
done
