#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tracking"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "tracking or outer" 2>&1 | tail -5
echo "== tracking warm"; timeout 300 python scripts/bench_tracking.py 2>&1 | tail -1 | tee gpurun_out/tracking_panda_warm.json
echo "== tracking cold"; timeout 300 python scripts/bench_tracking.py --cold 2>&1 | tail -1 | tee gpurun_out/tracking_panda_cold.json
echo "== tracking talos warm"; timeout 300 python scripts/bench_tracking.py --robot talos --batch 16384 --cpu-sample 1024 2>&1 | tail -1 | tee gpurun_out/tracking_talos_warm.json
