#!/bin/bash
# Round-2 state check in one gpurun call: GPU parity tests, the bench line (with the C3/C4/C5
# block), the launch list of a step.   bash tools/gpu_r02.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv 2>/dev/null | tail -40
