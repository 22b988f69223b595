#!/bin/bash
# development aid: A/B builds of the library on the GPU box (gpurun).  usage: ab_run.sh "<lib suffixes>" "<robots>"
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
LIBS=${1:-"v3 mig"}; ROBOTS=${2:-"panda,ur10,talos"}
for lib in $LIBS; do
  export LOIK_B200_LIB=$PWD/loik_b200/libloik_b200_$lib.so
  echo "== pytest $lib" ; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
  echo "== quick_perf $lib"; PIPE=1 DEPTHS=16,32 timeout 600 python scripts/quick_perf.py $ROBOTS 2>&1 | grep -v "^ *$" | tail -14
done
python scripts/solve_trace.py panda > gpurun_out/trace_panda.txt 2>&1
python scripts/solve_trace.py ur10 > gpurun_out/trace_ur10.txt 2>&1
