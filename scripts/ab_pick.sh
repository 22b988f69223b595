#!/bin/bash
# development aid (gpurun, 1 GPU, short on GPU-minutes): time the pipelined Panda solve of each library variant, install the
# fastest as the in-tree library ON THE BOX, then run the parity suite and the bench with it.  gpurun_out/ab_pick.txt
# records the choice; the source defaults are then set to match here.
mkdir -p gpurun_out; export CUDA_DEVICE_MAX_CONNECTIONS=32
{
best=default; bestv=0
for lib in ${1:-default}; do
  if [ "$lib" = default ]; then unset LOIK_B200_LIB; else export LOIK_B200_LIB=$PWD/loik_b200/libloik_b200_$lib.so; fi
  out=$(PIPE=1 DEPTHS=32 timeout 100 python scripts/quick_perf.py panda 2>&1 | grep -v "^ *$" | tail -3)
  echo "== $lib"; echo "$out"
  v=$(echo "$out" | grep "pipeline depth" | sed -E 's/.*, ([0-9.]+) M solves.*/\1/')
  if [ -n "$v" ] && awk "BEGIN{exit !($v > $bestv)}"; then best=$lib; bestv=$v; fi
done
unset LOIK_B200_LIB
echo "== chosen: $best ($bestv M solves/s)"
if [ "$best" != default ]; then cp loik_b200/libloik_b200_$best.so loik_b200/libloik_b200.so; fi
echo "== pytest ($best)"; timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench ($best)"; timeout 200 python bench.py > gpurun_out/bench_pick_panda.json 2> gpurun_out/bench_pick_panda.err; tail -c 900 gpurun_out/bench_pick_panda.json
} 2>&1 | tee gpurun_out/ab_pick.txt
