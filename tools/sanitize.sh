#!/bin/bash
# compute-sanitizer memcheck + racecheck over smoke() and tools/san_paths.py (run through gpurun);
# the summaries are kept under profiles/:   bash tools/sanitize.sh r02
tag=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_san_${tool}_smoke.log 2>&1
  compute-sanitizer --tool $tool --print-limit 20 python tools/san_paths.py > gpurun_out/${tag}_san_${tool}_paths.log 2>&1
done
tail -n 4 gpurun_out/${tag}_san_*.log
