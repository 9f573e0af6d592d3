#!/bin/bash
# one gpurun call of the development loop: parity tests of the reduction, phase profile, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "representations" 2>&1 | tail -15 > gpurun_out/dev_tests.log
cat gpurun_out/dev_tests.log
python tools/prof_fused.py > gpurun_out/dev_prof.log 2>&1; cat gpurun_out/dev_prof.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/dev_bench.log 2>&1; cat gpurun_out/dev_bench.log
python tools/one_push.py 5000 2 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dev.csv python tools/one_push.py 5000 2 >/dev/null 2>&1
grep -v "^==" gpurun_out/launches_dev.csv | python -c "
import csv,sys
for r in list(csv.reader(sys.stdin))[-5:]: print(r[4][:50], r[-1])
"
