#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2r
( timeout 900 python -m pytest tests -q -x -m gpu ) > ${O}_pytest.out 2>&1
echo "pytest rc=$?"; tail -2 ${O}_pytest.out
python tools/op_profile.py > ${O}_op_profile.txt 2> ${O}_op_profile.err; echo "op_profile rc=$?"; head -1 ${O}_op_profile.txt
grep "^d0" ${O}_op_profile.txt
