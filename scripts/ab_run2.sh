#!/bin/bash
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest"; python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for nw in 2 3 4; do echo "== talos NWARP=$nw"; LOIK_NWARP=$nw PIPE=1 DEPTHS=16 python scripts/quick_perf.py talos 2>&1 | grep -v "^ *$" | tail -3; done
echo "== panda depths"; PIPE=1 DEPTHS=16,24,32,48 python scripts/quick_perf.py panda 2>&1 | grep pipeline
echo "== ur10 depths"; PIPE=1 DEPTHS=8,16,24 python scripts/quick_perf.py ur10 2>&1 | grep pipeline
echo "== talos depths"; PIPE=1 DEPTHS=24,32,48 python scripts/quick_perf.py talos 2>&1 | grep pipeline
