#!/bin/bash
# launch list + one full ncu capture of the streamed-reduction kernels on one named shape
#   bash tools/gpu_sweep_prof.sh <tag> <cfg> <frames> <tile>
tag=$1; cfg=${2:-C4}; fr=${3:-32}; tile=${4:-8}
mkdir -p gpurun_out
python tools/bench_configs.py --configs $cfg --frames $fr --tile $tile --reps 1 > gpurun_out/${tag}_cfg_${cfg}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launch_${cfg}.csv \
    python tools/bench_configs.py --configs $cfg --frames $fr --tile $tile --reps 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launch_${cfg}.csv | grep -v "at::" | head -12
ncu --set full --clock-control none --import-source on -k "regex:k_sweep_l|k_sweep_bt" -c 2 -o gpurun_out/${tag}_sweep_${cfg} \
    python tools/bench_configs.py --configs $cfg --frames $fr --tile $tile --reps 1 > gpurun_out/${tag}_ncu_${cfg}.log 2>&1
