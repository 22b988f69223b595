#!/bin/bash
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== perf"; PIPE=1 DEPTHS=32 timeout 300 python scripts/quick_perf.py panda,talos,talos_ff 2>&1 | grep -v "^ *$" | tail -9
