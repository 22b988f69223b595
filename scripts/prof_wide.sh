set -x
mkdir -p gpurun_out/wide
for g in 1 4; do
  BATCH=4 GPI=$g timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_iterate_lane -s 2 -c 1 -o gpurun_out/wide/talos_g$g -f python scripts/lane_prof.py talos 0 20 > gpurun_out/wide/log_g$g.txt 2>&1
done
ls -la gpurun_out/wide
