#!/bin/bash
# Produces the measured evidence kept under profiles/ (run through gpurun; summaries are made
# afterwards on the CPU box with tools/ncu_summary.py):  bash tools/gpu_profiles.sh r02
tag=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference.json 2> gpurun_out/${tag}_reference.err
# every launch with its device time (cold-cache, serialised: compare shares)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches.log 2>&1
# one full capture per kernel of a step (warm-up: 1 + 3 reductions of 5 kernels, 3 x (2 msd + lut + rdf))
ncu --set full --clock-control none --import-source on -k "regex:k_chunk_box|k_img3|k_scan3|k_evsort|k_frames4|k_msd_pairs|k_rdf5" -s 36 -c 9 \
    -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_full.log 2>&1
# the other named shapes: launch lists and one full capture of the streamed-reduction kernels
for spec in C3:64:4 C4:32:8 C5:48:4; do
  IFS=: read cfg fr tile <<< "$spec"
  python tools/bench_configs.py --configs $cfg --frames $fr --tile $tile --reps 3 > gpurun_out/${tag}_cfg_${cfg}.log 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launch_${cfg}.csv \
      python tools/bench_configs.py --configs $cfg --frames $fr --tile $tile --reps 1 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k "regex:k_sweep_p|k_sweep_s|k_sweep_l|k_sweep_bt|k_zpart|k_labels" -c 6 \
      -o gpurun_out/${tag}_sweep_${cfg} python tools/bench_configs.py --configs $cfg --frames $fr --tile $tile --reps 1 > gpurun_out/${tag}_ncu_${cfg}.log 2>&1
done
cat gpurun_out/${tag}_bench.json
cat gpurun_out/${tag}_reference.json
