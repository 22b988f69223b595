#!/bin/bash
# Everything the round's profiles/ are made from, in one gpurun call (1 GPU): tests, bench lines, ncu launch list + full captures.
TAG=${1:-r1}
mkdir -p gpurun_out
echo "== pytest -m gpu"; python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench (default = panda)"; python bench.py > gpurun_out/bench_${TAG}_panda.json 2> gpurun_out/bench_${TAG}_panda.err; tail -c 600 gpurun_out/bench_${TAG}_panda.json
for w in ur10 talos; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err; done
echo "== reference arm"; python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>&1
echo "== tracking / single instance"
python scripts/bench_tracking.py > gpurun_out/tracking_${TAG}_panda_warm.json 2>/dev/null
python scripts/bench_tracking.py --cold > gpurun_out/tracking_${TAG}_panda_cold.json 2>/dev/null
python scripts/bench_tracking.py --robot talos --batch 16384 --cpu-sample 1024 > gpurun_out/tracking_${TAG}_talos_warm.json 2>/dev/null
python scripts/single_instance.py > gpurun_out/single_instance_${TAG}.json 2>/dev/null
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --pipeline 1 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
echo "== ncu full"
for r in panda ur10 talos; do
  ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 10 -c 2 -f -o gpurun_out/prof_${TAG}_$r \
      python scripts/quick_perf.py $r > gpurun_out/prof_${TAG}_$r.log 2>&1
done
ls -la gpurun_out | tail -20
