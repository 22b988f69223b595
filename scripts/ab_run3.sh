#!/bin/bash
mkdir -p gpurun_out
python scripts/solve_trace.py panda > gpurun_out/trace_panda.txt 2>&1
python scripts/solve_trace.py ur10 > gpurun_out/trace_ur10.txt 2>&1
DEPTH=16 python scripts/timeline.py ur10 > gpurun_out/timeline_ur10.txt 2>&1
DEPTH=32 python scripts/timeline.py panda > gpurun_out/timeline_panda32.txt 2>&1
tail -5 gpurun_out/timeline_ur10.txt gpurun_out/timeline_panda32.txt
