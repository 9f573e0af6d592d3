#!/bin/bash
# one gpurun call of the development loop: GPU parity tests, phase profile, bench, launch list
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/dev_tests.log
cat gpurun_out/dev_tests.log
python tools/prof_fused.py > gpurun_out/dev_prof.log 2>&1; cat gpurun_out/dev_prof.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/dev_bench.log 2>&1; cat gpurun_out/dev_bench.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dev.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline >/dev/null 2>&1
grep -v "^==" gpurun_out/launches_dev.csv | python -c "
import csv,sys,collections
rows=list(csv.reader(sys.stdin))
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[4][:60],[]).append(float(r[-1]))
    except Exception: pass
for k,v in agg.items(): print('%-62s n=%4d  last=%10.1f  min=%10.1f' % (k,len(v),v[-1],min(v)))
"
