#!/bin/bash
# Everything the round's profiles/ are made from, in one gpurun call (1 GPU): tests, bench lines, ncu launch list + full captures
# of the benched build, DRAM traffic of the roofline kernel (-> profiles/traffic.json via scripts/summarize_profiles.py).
TAG=${1:-r2}
mkdir -p gpurun_out
echo "== pytest -m gpu"; python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench (default: panda headline + ur10 / talos sub-records)"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 300 gpurun_out/bench_${TAG}.json
echo "== bench as the driver runs it"; python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_driver.json 2> gpurun_out/bench_${TAG}_driver.err
echo "== reference arm"; python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>&1
echo "== tracking / single instance"
python scripts/bench_tracking.py > gpurun_out/tracking_${TAG}_panda_warm.json 2>/dev/null
python scripts/bench_tracking.py --robot talos --batch 16384 --cpu-sample 1024 > gpurun_out/tracking_${TAG}_talos_warm.json 2>/dev/null
python scripts/single_instance.py > gpurun_out/single_instance_${TAG}.json 2>/dev/null
echo "== schedule sweep (switch point of the lane-parallel kernel)"
LANE_AFTERS=-1,10,32 DEPTHS=1,4,10,32 python scripts/lane_perf.py panda 2>&1 | tail -3 > gpurun_out/lane_sweep_${TAG}.txt
LANE_AFTERS=-1,32 DEPTHS=1,4,16 python scripts/lane_perf.py ur10 2>&1 | tail -2 >> gpurun_out/lane_sweep_${TAG}.txt
LANE_AFTERS=-1,8,32 DEPTHS=1,4,16 python scripts/lane_perf.py talos 2>&1 | tail -3 >> gpurun_out/lane_sweep_${TAG}.txt
python scripts/lane_latency.py panda,ur10,talos 2>&1 | tail -9 > gpurun_out/lane_latency_${TAG}.txt
cat gpurun_out/lane_sweep_${TAG}.txt
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --workload panda --steps 2 --warmup 3 --pipeline 1 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
echo "== ncu full: the roofline kernel (k_iterate, one dense iteration per launch) and the lane-parallel kernel"
# (scripts/lane_prof.py <robot> -1 1: three fixed-iteration launches of ONE dense iteration each, then a Solve(); the second
#  and third of those launches are captured)
for r in panda ur10 talos; do
  ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 1 -c 2 -f -o gpurun_out/prof_${TAG}_$r \
      python scripts/lane_prof.py $r -1 1 > gpurun_out/prof_${TAG}_$r.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_iterate_lane -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_lane_panda \
    python scripts/lane_prof.py panda 0 4 > gpurun_out/prof_${TAG}_lane_panda.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_iterate_lane -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_lane_talos \
    python scripts/lane_prof.py talos 0 4 > gpurun_out/prof_${TAG}_lane_talos.log 2>&1
# the summaries are made here, next to the captures (five .ncu-rep files exceed what gpurun brings back); the two Panda captures travel too
PROFILES_OUT=gpurun_out/profiles_out python scripts/summarize_profiles.py ${TAG} > gpurun_out/summarize_${TAG}.log 2>&1
rm -f gpurun_out/prof_${TAG}_ur10.ncu-rep gpurun_out/prof_${TAG}_talos.ncu-rep gpurun_out/prof_${TAG}_lane_talos.ncu-rep
ls -la gpurun_out | tail -24
