#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2v
( timeout 900 python -m pytest tests -q -x -m gpu ) > ${O}_pytest.out 2>&1
echo "pytest rc=$?"; tail -2 ${O}_pytest.out
python tools/op_profile.py > ${O}_op_profile.txt 2> ${O}_op_profile.err; echo "op_profile rc=$?"; head -1 ${O}_op_profile.txt
grep "^attn" ${O}_op_profile.txt
( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 ) > ${O}_bench.out 2> ${O}_bench.err
echo "bench rc=$?"; grep '^{' ${O}_bench.out | cut -c1-200
