#!/bin/bash
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest"; python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "LOIK_SEG_AFTER=8" "LOIK_SEG_AFTER=16" "LOIK_SEG_AFTER=32" "LOIK_SEG_AFTER=16 LOIK_HI_AFTER=16"; do echo "== talos $cfg"; env $cfg PIPE=1 DEPTHS=8,16,32 python scripts/quick_perf.py talos 2>&1 | grep -v "^ *$" | tail -5; done
echo "== talos_ff"; PIPE=1 DEPTHS=16 python scripts/quick_perf.py talos_ff 2>&1 | grep -v "^ *$" | tail -3
