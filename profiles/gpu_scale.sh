#!/bin/bash
# Scaling bench lines only (ONE gpurun --gpus N command):  bash profiles/gpu_scale.sh <tag> <N> <config> [extra bench args]
set -u
tag=$1; n=$2; cfg=$3; shift; shift; shift
out=gpurun_out; mkdir -p $out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --config $cfg "$@" \
    > $out/${tag}_bench_${cfg}_n$n.json 2> $out/${tag}_bench_${cfg}_n$n.err; echo "bench $cfg N=$n rc=$?"; grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_bench_${cfg}_n$n.err | tail -c 1500
