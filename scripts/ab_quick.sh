#!/bin/bash
# development aid: quick A/B of library variants on the GPU box (gpurun, 1 GPU): parity suite on the variant named
# first, then dense-iteration / single-solve / pipelined timings of each.
# usage: ab_quick.sh "<lib suffixes, 'default' = the in-tree build>" "<robots>"
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
setlib() { if [ "$1" = default ]; then unset LOIK_B200_LIB; else export LOIK_B200_LIB=$PWD/loik_b200/libloik_b200_$1.so; fi; }
{
first=$(echo ${1:-default} | awk '{print $1}')
setlib $first
echo "== pytest $first"; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for lib in ${1:-default}; do
  setlib $lib
  echo "== quick_perf $lib"; PIPE=1 DEPTHS=32 timeout 200 python scripts/quick_perf.py ${2:-panda} 2>&1 | grep -v "^ *$" | tail -12
done
} 2>&1 | tee gpurun_out/ab_quick.txt
