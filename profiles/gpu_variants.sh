#!/bin/bash
# A/B of tuning knobs (ONE gpurun command): bash profiles/gpu_variants.sh <tag> <config> "ENV=val ..." "ENV=val ..." ...
# every quoted group is one variant: a short bench line (no extras) with those environment variables set.
set -u
tag=$1; cfg=$2; shift; shift
out=gpurun_out; mkdir -p $out
i=0
for v in "$@"; do
  i=$((i + 1))
  env $v timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_var${i}_$cfg.json 2> $out/${tag}_var${i}_$cfg.err
  echo "variant $i [$v] rc=$?: $(python - <<PY
import json
j=json.loads([l for l in open('$out/${tag}_var${i}_$cfg.json') if l.startswith('{')][-1])
t=j['stage_ms_rank0']
print('ms/step %.3f  prep %.3f  read_kernel %.3f  walk_kernel %.3f  filter %.3f  e2e %.3f' % (j['ms_per_step'], t['graph_prep_ms'], t['read_kernel_ms'], t['walk_kernel_ms'], t['filter_ms'], j['e2e']['ms_per_step']))
PY
)"
done
