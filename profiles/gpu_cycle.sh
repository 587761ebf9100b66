#!/bin/bash
# One measurement cycle on a GPU box, meant to be the command of ONE gpurun call:
#   gpurun --timeout 900 -- 'bash profiles/gpu_cycle.sh v18 [kernel-regex-for-ncu-full]'
# 1. pytest -m gpu (stops at the first failure)   2. bench.py (N=1, with the CPU baseline)   3. ncu launch list of a short bench run
# 4. optional: ncu --set full of the kernels matching the regex (second resident step).  Everything lands in gpurun_out/<tag>_*;
# `python profiles/collect.py <tag>` then files the summaries under profiles/ (run it back in the dev container).
set -u
tag=${1:-cycle}; regex=${2:-}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_list.log 2>&1; echo "launch list rc=$?"
if [ -n "$regex" ]; then
    n=$(echo "$regex" | tr '|' '\n' | wc -l)
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $((3 * n)) -c $n -o $out/${tag}_full -f \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
