#!/bin/bash
# Profiling cycle (ONE gpurun command): bash profiles/gpu_prof.sh <tag> <config> [kernel regex for ncu --set full]
# bench line without extras, ncu launch list of the same short command, optional full capture of the named kernels.
set -u
tag=${1:-prof}; cfg=${2:-c4}; regex=${3:-}
out=gpurun_out; mkdir -p $out
timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_profbench_$cfg.json 2> $out/${tag}_profbench_$cfg.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches_$cfg.csv \
    python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_ncu_list.log 2>&1; echo "launch list rc=$?"
if [ -n "$regex" ]; then
    n=$(echo "$regex" | tr '|' '\n' | wc -l)
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $((3 * n)) -c $n -o $out/${tag}_full_$cfg -f \
        python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
