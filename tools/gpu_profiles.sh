#!/bin/bash
# Produces the measured evidence kept under profiles/ (run through gpurun; summaries are made
# afterwards on the CPU box with tools/ncu_summary.py):  bash tools/gpu_profiles.sh r01
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference.json 2> gpurun_out/${tag}_reference.err
# every launch with its device time (cold-cache, serialised: compare shares)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# one full capture per kernel of a step (warm-up: 1 + 3 reductions, 3 x (2 msd + lut + rdf))
ncu --set full --clock-control none --import-source on -k "regex:k_img3|k_scan3|k_frames3|k_msd_pairs|k_rdf5" -s 24 -c 7 \
    -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/${tag}_bench.json
cat gpurun_out/${tag}_reference.json
