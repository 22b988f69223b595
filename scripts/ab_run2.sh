#!/bin/bash
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
for mb in 4 5; do echo "== MINB=$mb"; LOIK_MINB=$mb PIPE=1 DEPTHS=32 python scripts/quick_perf.py panda,ur10 2>&1 | grep -v "^ *$" | tail -6; done
