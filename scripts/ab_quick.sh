#!/bin/bash
# development aid: tests + smoke on the default build, then a quick A/B of library variants (gpurun, 1 GPU).
# usage: ab_quick.sh "<lib suffixes, '' = default>" "<robots>"
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest"; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for lib in ${1:-"default"}; do
  if [ "$lib" = default ]; then unset LOIK_B200_LIB; else export LOIK_B200_LIB=$PWD/loik_b200/libloik_b200_$lib.so; fi
  echo "== quick_perf $lib"; PIPE=1 DEPTHS=32 timeout 200 python scripts/quick_perf.py ${2:-panda} 2>&1 | grep -v "^ *$" | tail -8
done 2>&1 | tee gpurun_out/ab_quick.txt
