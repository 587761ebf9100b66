#!/bin/bash
# Multi-GPU cycle (ONE gpurun --gpus N command):  bash profiles/gpu_multi.sh <tag> <N> [configs...]
# the N-rank parity tests, then one torchrun bench line per config.
set -u
tag=${1:-multi}; n=${2:-2}; shift; shift
cfgs=${@:-c4}
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests/test_multi.py tests/test_gpu_dropin.py -m gpu -x -q -k "multi_gpu or several_gpus" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/${tag}_pytest.log
for c in $cfgs; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config $c \
      > $out/${tag}_bench_${c}_n$n.json 2> $out/${tag}_bench_${c}_n$n.err; echo "bench $c N=$n rc=$?"; tail -c 1500 $out/${tag}_bench_${c}_n$n.err
done
