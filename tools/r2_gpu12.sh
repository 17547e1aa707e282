#!/bin/bash
# Round-2 GPU session 12: sk epilogue straight to global memory from 16x256b TMEM fragments (no R ring / staging / TMA stores).
mkdir -p gpurun_out
O=gpurun_out/r2l
( timeout 900 python -m pytest tests -q -x -m gpu ) > ${O}_pytest.out 2>&1
echo "pytest rc=$?"; tail -5 ${O}_pytest.out | cut -c1-300
python tools/op_profile.py > ${O}_op_profile.txt 2> ${O}_op_profile.err; echo "op_profile rc=$?"; head -1 ${O}_op_profile.txt
grep -A45 "by (kind" ${O}_op_profile.txt | grep "^sk"
( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 ) > ${O}_bench.out 2> ${O}_bench.err
echo "bench rc=$?"; grep '^{' ${O}_bench.out | cut -c1-200
python tools/sk_timeline.py --ops 6:conv1,4:out,7:inject > ${O}_timeline.txt 2>&1
