#!/bin/bash
# Per-kernel launch lists of the C3 / C4 / C5 stage benchmarks (tools/bench_configs.py), one ncu
# launch-list pass per config:  bash tools/gpu_cfg_launches.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
for spec in C3:64:4 C4:32:8 C5:48:4; do
  IFS=: read cfg fr tile <<< "$spec"
  python tools/bench_configs.py --configs $cfg --frames $fr --tile $tile --reps 1 > gpurun_out/${tag}_cfg_${cfg}.log 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launch_${cfg}.csv \
      python tools/bench_configs.py --configs $cfg --frames $fr --tile $tile --reps 1 > /dev/null 2>&1
  echo "== $cfg"; cat gpurun_out/${tag}_cfg_${cfg}.log | cut -c1-400
  python tools/launch_summary.py gpurun_out/${tag}_launch_${cfg}.csv | grep -v "at::"
done
