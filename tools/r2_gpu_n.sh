#!/bin/bash
# Multi-GPU bench exactly as the driver launches it: N from $1
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r2n${N}
( timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 ) > ${O}_bench.out 2> ${O}_bench.err
echo "== bench N=$N rc=$?"
grep '^{' ${O}_bench.out | tail -c 3000
grep -v "^\s*$" ${O}_bench.err | grep -v "Warning\|warn" | tail -n 15 | cut -c1-300
( timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) > ${O}_ref.out 2> ${O}_ref.err
echo "== reference N=$N rc=$?"
grep '^{' ${O}_ref.out | cut -c1-400
