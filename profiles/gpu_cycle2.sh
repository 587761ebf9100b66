#!/bin/bash
# Round-2 measurement cycle (ONE gpurun command):  bash profiles/gpu_cycle2.sh <tag> [configs...]
# pytest -m gpu, then one bench line per listed config (default: c2 c4), then the ncu launch list of a short c4 run.
set -u
tag=${1:-cycle}; shift
cfgs=${@:-c2 c4}
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
for c in $cfgs; do
  timeout 900 python bench.py --config $c > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err; echo "bench $c rc=$?"; tail -c 600 $out/${tag}_bench_$c.err
done
