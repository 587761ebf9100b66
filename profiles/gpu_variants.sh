#!/bin/bash
# A/B of the walk kernel register budgets (ONE gpurun command): bash profiles/gpu_variants.sh <tag> <config>
set -u
tag=$1; cfg=${2:-c2}; out=gpurun_out; mkdir -p $out
for v in 6 7 8; do
  PHI_GPU_WALK_CTAS=$v timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_walkctas${v}_$cfg.json 2> $out/${tag}_walkctas${v}_$cfg.err; echo "ctas=$v rc=$?"
done
